// Issue-rate probe for sm_100a: scalar FP32 (FFMA / FMUL / FADD) against the packed two-lane forms
// (fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2 -> FFMA2 / FMUL2 / FADD2).  Prints warp instructions per cycle and SM
// and FP32 lane-operations per cycle and SM for 4..32 resident warps per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
template <int MODE>
__global__ void probe(float* out, float a, float b, long long* cyc) {
  float x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = x[i] + 0.5f; }
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("fma.rn.f32 %0, %0, %2, %3; fma.rn.f32 %1, %1, %2, %3;" : "+f"(x[i]), "+f"(y[i]) : "f"(a), "f"(b)); }
      if (MODE == 1) { asm volatile("{.reg .b64 r, s, t; mov.b64 r, {%0,%1}; mov.b64 s, {%2,%2}; mov.b64 t, {%3,%3}; fma.rn.f32x2 r, r, s, t; mov.b64 {%0,%1}, r;}" : "+f"(x[i]), "+f"(y[i]) : "f"(a), "f"(b)); }
      if (MODE == 2) { asm volatile("mul.rn.f32 %0, %0, %2; mul.rn.f32 %1, %1, %2;" : "+f"(x[i]), "+f"(y[i]) : "f"(a)); }
      if (MODE == 3) { asm volatile("{.reg .b64 r, s; mov.b64 r, {%0,%1}; mov.b64 s, {%2,%2}; mul.rn.f32x2 r, r, s; mov.b64 {%0,%1}, r;}" : "+f"(x[i]), "+f"(y[i]) : "f"(a)); }
      if (MODE == 4) { asm volatile("add.rn.f32 %0, %0, %2; add.rn.f32 %1, %1, %2;" : "+f"(x[i]), "+f"(y[i]) : "f"(b)); }
      if (MODE == 5) { asm volatile("{.reg .b64 r, s; mov.b64 r, {%0,%1}; mov.b64 s, {%2,%2}; add.rn.f32x2 r, r, s; mov.b64 {%0,%1}, r;}" : "+f"(x[i]), "+f"(y[i]) : "f"(b)); }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int lanes_per_instr, int instr_per_slot) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  for (int warps = 4; warps <= 32; warps *= 2) {
    probe<MODE><<<sms, warps * 32>>>(out, 1.0000001f, 1e-9f, cyc);
    probe<MODE><<<sms, warps * 32>>>(out, 1.0000001f, 1e-9f, cyc);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += (double)h[i]; c /= sms;
    const double winstr = (double)ITER * 8 * instr_per_slot * warps;           // warp instructions per SM
    printf("%-6s warps/SM %2d  cycles %9.0f  warp-instr/cycle/SM %.3f  fp32 lane-ops/cycle/SM %.1f\n", name, warps, c, winstr / c,
           winstr * 32 * lanes_per_instr / c);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 1, 2); run<1>("FFMA2", 2, 1);
  run<2>("FMUL", 1, 2); run<3>("FMUL2", 2, 1);
  run<4>("FADD", 1, 2); run<5>("FADD2", 2, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
