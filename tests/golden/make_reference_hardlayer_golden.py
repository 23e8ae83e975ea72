"""Golden vectors for HardLayer made by the REFERENCE'S OWN class (Hard_Layer_Subclassed.py:21-260, cut out by AST) executed
through the torch-backed TensorFlow stand-in: the example's configuration (no rbf, no rectifier, identity activations,
identity nonormalize_func, norm_limits [-1, 1], per-cell trainable kernel_exponent).  Values and the cotangents torch
autograd delivers for the network output, for kernel_exponent and for the time input (time inputs > t_lo: tf.pow's and torch.pow's
gradients coincide there).   Output: tests/golden/reference_hardlayer.npz
"""
import ast
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_torch_shim as tf          # noqa: E402

REF = "/root/reference/Hard_Layer_Subclassed.py"


def main():
    src = open(REF).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "HardLayer")
    ns = {"tf": tf, "np": np, "get_configuration": lambda *a, **k: {"dew_point": 4048.49},
          "DEFAULT_GENERAL_CONFIG": {"fluid_type": "DG"}, "DEFAULT_RESERVOIR_CONFIG": {"initialization": {"Pi": 5000.0, "Pa": 14.7}}}
    exec(textwrap.dedent(ast.get_source_segment(src, node)), ns)
    rng = np.random.default_rng(5600)
    B, D, H, W = 5, 2, 4, 6
    layer = ns["HardLayer"](norm_limits=[-1, 1], init_value=5000.0,
                            kernel_exponent_config={"initial_value": (0.5,), "trainable": True, "min_value": 0.1, "max_value": 1.0})
    layer.build([(B, D, H, W, 1), (B, D, H, W, 1)])
    expo = (0.1 + 0.85 * rng.random((D, H, W, 1))).astype(np.float32)
    layer.kernel_exponent = torch.as_tensor(expo).requires_grad_(True)
    tn = np.asarray([-1.0, -0.6, 0.0, 0.45, 1.0], np.float32)
    tn_t = torch.as_tensor(tn).requires_grad_(True)          # the layer's time input is differentiable (physics_loss.py:105-111)
    time = tn_t.view(B, 1, 1, 1, 1).expand(B, D, H, W, 1)
    prop = torch.zeros(B, D, H, W, 1)
    y = torch.as_tensor((600.0 * rng.random((B, D, H, W, 1))).astype(np.float32)).requires_grad_(True)
    out = layer([[time, prop], y])
    wgt = torch.as_tensor(rng.standard_normal((B, D, H, W, 1)).astype(np.float32))
    # cotangents over the samples with alpha_t > 0 only (at alpha_t = 0 torch.pow's exponent gradient is nan, tf's is 0)
    sel = torch.as_tensor((tn > -1.0).astype(np.float32)).view(B, 1, 1, 1, 1)
    gy, ge, gt = torch.autograd.grad((out * wgt * sel).sum(), [y, layer.kernel_exponent, tn_t])
    np.savez_compressed(os.path.join(HERE, "reference_hardlayer.npz"), tn=tn, expo=expo[..., 0], y=y.detach().numpy()[..., 0],
                        out=out.detach().numpy()[..., 0], wgt=(wgt * sel).numpy()[..., 0], gy=gy.numpy()[..., 0], gexpo=ge.numpy()[..., 0],
                        gtn=gt.numpy())
    print("wrote reference_hardlayer.npz", out.shape)
    options_case(src, node)


def options_case(src, node):
    """The layer's non-default options together (Hard_Layer_Subclassed.py:41-45, 168-187, 228-246): use_rbf (a
    Dense(1, sigmoid) on the property channel), the gas-condensate rectifier on a third input, an activation on the
    kernel exponent and one on the network output.  Output: tests/golden/reference_hardlayer_opts.npz"""
    ns = {"tf": tf, "np": np, "get_configuration": lambda *a, **k: {"dew_point": 4048.49},
          "DEFAULT_GENERAL_CONFIG": {"fluid_type": "GC"}, "DEFAULT_RESERVOIR_CONFIG": {"initialization": {"Pi": 5000.0, "Pa": 14.7}}}
    exec(textwrap.dedent(ast.get_source_segment(src, node)), ns)
    rng = np.random.default_rng(5601)
    B, D, H, W = 4, 2, 3, 5
    layer = ns["HardLayer"](norm_limits=[-1, 1], init_value=5000.0,
                            kernel_exponent_config={"initial_value": (0.5,), "trainable": True, "min_value": 0.1, "max_value": 1.0},
                            use_rbf=True, rbf_config={"output_dim": 25, "activation": "sigmoid"}, rectifier=tf.nn.relu,
                            kernel_activation=[tf.nn.sigmoid], input_activation=tf.nn.tanh)
    layer.build([(B, D, H, W, 1), (B, D, H, W, 1)])
    expo = (0.1 + 0.85 * rng.random((D, H, W, 1))).astype(np.float32)
    layer.kernel_exponent = torch.as_tensor(expo).requires_grad_(True)
    layer.rbf_dense.kernel = torch.tensor([[0.7]], dtype=torch.float32, requires_grad=True)
    layer.rbf_dense.bias = torch.tensor([-0.2], dtype=torch.float32, requires_grad=True)
    tn = np.asarray([-0.6, 0.0, 0.45, 1.0], np.float32)
    tn_t = torch.as_tensor(tn).requires_grad_(True)
    time = tn_t.view(B, 1, 1, 1, 1).expand(B, D, H, W, 1)
    prop = torch.as_tensor(rng.uniform(-1, 1, (B, D, H, W, 1)).astype(np.float32))
    y = torch.as_tensor(rng.uniform(-1.5, 1.5, (B, D, H, W, 1)).astype(np.float32)).requires_grad_(True)
    rect = torch.as_tensor(rng.uniform(3000.0, 4600.0, (B, D, H, W, 1)).astype(np.float32)).requires_grad_(True)   # both sides of the dew point
    out = layer([[time, prop], y, rect])
    wgt = torch.as_tensor(rng.standard_normal((B, D, H, W, 1)).astype(np.float32))
    gy, ge, gt, gk, gb, gr = torch.autograd.grad((out * wgt).sum(), [y, layer.kernel_exponent, tn_t, layer.rbf_dense.kernel,
                                                                     layer.rbf_dense.bias, rect])
    np.savez_compressed(os.path.join(HERE, "reference_hardlayer_opts.npz"), tn=tn, expo=expo[..., 0], y=y.detach().numpy()[..., 0],
                        prop=prop.numpy()[..., 0], rect=rect.detach().numpy()[..., 0], out=out.detach().numpy()[..., 0],
                        wgt=wgt.numpy()[..., 0], gy=gy.numpy()[..., 0], gexpo=ge.numpy()[..., 0], gtn=gt.numpy(),
                        gkernel=gk.numpy(), gbias=gb.numpy(), grect=gr.numpy()[..., 0], kernel=np.float32(0.7), bias=np.float32(-0.2),
                        pdew=np.float32(4048.49), pmin=np.float32(14.7), init_value=np.float32(5000.0))
    print("wrote reference_hardlayer_opts.npz", out.shape, "rectifier active on", int((rect.detach() < 4048.49).sum()), "of", rect.numel())


if __name__ == "__main__":
    main()
