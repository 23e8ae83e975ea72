// TensorFlow custom op binding libsrm_physics.so (include/srm_physics.h) into the reference.
//
// NOT BUILT IN THIS REPOSITORY'S CI: TensorFlow is not installable in the build container
// (SURVEY.md F3).  On a host with TensorFlow >= 2.11:
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++17 -shared -fPIC tf_op/srm_physics_op.cc -o tf_op/srm_physics_op.so \
//       -Iinclude $TF_CFLAGS $TF_LFLAGS -L<pkg dir> -lsrm_physics -DGOOGLE_CUDA=1
//
// Two ops (forward, backward) sharing an SrmHandle owned by a resource created from Python;
// the Python side registers SrmPhysicsBackward as the gradient of SrmPhysicsForward.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/common_runtime/gpu/gpu_event_mgr.h"
#include "tensorflow/core/platform/stream_executor.h"

#include "srm_physics.h"

using namespace tensorflow;

REGISTER_OP("SrmPhysicsForward")
    .Attr("handle: int")          // SrmHandle* smuggled as int64 (created by srm_create through ctypes)
    .Input("kx: float")           // (R,D,H,W)
    .Input("sample_real: int32")  // (B,)
    .Input("p0: float")           // (B,D,H,W)
    .Input("p1: float")
    .Input("dt1: float")          // (B,)
    .Input("dt2: float")
    .Input("t1: float")
    .Output("terms: float")       // (2,8)
    .Output("workspace: uint8")   // kept alive for the backward op
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->MakeShape({2, SRM_N_TERMS}));
      c->set_output(1, c->UnknownShapeOfRank(1));
      return OkStatus();
    });

REGISTER_OP("SrmPhysicsBackward")
    .Attr("handle: int")
    .Input("kx: float").Input("sample_real: int32").Input("p0: float").Input("p1: float")
    .Input("dt1: float").Input("dt2: float").Input("t1: float")
    .Input("dterms: float")       // (8,) upstream gradient of terms[0]
    .Input("workspace: uint8")
    .Output("gp0: float").Output("gp1: float").Output("gdt1: float").Output("gdt2: float")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(2)); c->set_output(1, c->input(3));
      c->set_output(2, c->input(4)); c->set_output(3, c->input(5));
      return OkStatus();
    });

namespace {
void* CudaStream(OpKernelContext* ctx) {
  return reinterpret_cast<void*>(ctx->op_device_context()->stream()->platform_specific_handle().stream);
}
}  // namespace

class SrmPhysicsForwardOp : public OpKernel {
 public:
  explicit SrmPhysicsForwardOp(OpKernelConstruction* c) : OpKernel(c) {
    int64_t h; OP_REQUIRES_OK(c, c->GetAttr("handle", &h)); h_ = reinterpret_cast<SrmHandle*>(h);
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &kx = ctx->input(0), &sr = ctx->input(1), &p0 = ctx->input(2), &p1 = ctx->input(3),
                 &dt1 = ctx->input(4), &dt2 = ctx->input(5), &t1 = ctx->input(6);
    const int32_t B = p0.dim_size(0), R = kx.dim_size(0);
    Tensor *terms, *ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({2, SRM_N_TERMS}), &terms));
    const int64_t wsb = (int64_t)srm_workspace_bytes(h_, B, R, SRM_FLAG_SAVE_FOR_BACKWARD);
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({wsb}), &ws));
    int rc = srm_forward(h_, B, R, kx.flat<float>().data(), sr.flat<int32>().data(), p0.flat<float>().data(),
                         p1.flat<float>().data(), dt1.flat<float>().data(), dt2.flat<float>().data(),
                         t1.flat<float>().data(), terms->flat<float>().data(), nullptr, nullptr, nullptr,
                         ws->flat<uint8>().data(), (size_t)wsb, SRM_FLAG_SAVE_FOR_BACKWARD, CudaStream(ctx));
    OP_REQUIRES(ctx, rc == SRM_OK, errors::Internal("srm_forward: ", srm_last_error()));
  }
 private:
  SrmHandle* h_;
};

class SrmPhysicsBackwardOp : public OpKernel {
 public:
  explicit SrmPhysicsBackwardOp(OpKernelConstruction* c) : OpKernel(c) {
    int64_t h; OP_REQUIRES_OK(c, c->GetAttr("handle", &h)); h_ = reinterpret_cast<SrmHandle*>(h);
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &kx = ctx->input(0), &sr = ctx->input(1), &p0 = ctx->input(2), &p1 = ctx->input(3),
                 &dt1 = ctx->input(4), &dt2 = ctx->input(5), &t1 = ctx->input(6), &dterms = ctx->input(7),
                 &ws = ctx->input(8);
    const int32_t B = p0.dim_size(0), R = kx.dim_size(0);
    Tensor *gp0, *gp1, *gd1, *gd2;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, p0.shape(), &gp0));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, p1.shape(), &gp1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, dt1.shape(), &gd1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, dt2.shape(), &gd2));
    int rc = srm_backward(h_, B, R, kx.flat<float>().data(), sr.flat<int32>().data(), p0.flat<float>().data(),
                          p1.flat<float>().data(), dt1.flat<float>().data(), dt2.flat<float>().data(),
                          t1.flat<float>().data(), dterms.flat<float>().data(), gp0->flat<float>().data(),
                          gp1->flat<float>().data(), gd1->flat<float>().data(), gd2->flat<float>().data(),
                          const_cast<uint8*>(ws.flat<uint8>().data()), (size_t)ws.NumElements(), 0, CudaStream(ctx));
    OP_REQUIRES(ctx, rc == SRM_OK, errors::Internal("srm_backward: ", srm_last_error()));
  }
 private:
  SrmHandle* h_;
};

REGISTER_KERNEL_BUILDER(Name("SrmPhysicsForward").Device(DEVICE_GPU), SrmPhysicsForwardOp);
REGISTER_KERNEL_BUILDER(Name("SrmPhysicsBackward").Device(DEVICE_GPU), SrmPhysicsBackwardOp);
