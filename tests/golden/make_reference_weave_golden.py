"""Golden vectors for the feature-tensor construction made by the REFERENCE'S OWN code: `weave_tensors`
(data_processing/data_processing_utils.py:90-223) is cut out by AST and executed on the tensor list the data pipeline
hands it (srm_data_processing.py:363-403: permx (K, Nz, Ny, Nx), time (T, 1), x / y / z (1, Nz, Ny, Nx)) with
flatten_first_axes=True, followed by DataSummary.normalize (data_processing_utils.py:979-1063, 'lnk-linear-scaling':
linear rows 0..3, logarithmic permeability row 4) -- the (K*T, Nz, Ny, Nx, 5) tensor with channels [z, y, x, t, k] in
[-1, 1] that every training step consumes.  TensorFlow is replaced by the torch-backed stand-in of this directory.

Output: tests/golden/reference_weave.npz        python tests/golden/make_reference_weave_golden.py
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
from make_reference_wells_golden import build_class          # noqa: E402

REF = "/root/reference/data_processing/data_processing_utils.py"


def reference_function(name):
    src = open(REF).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    return textwrap.dedent(ast.get_source_segment(src, fn))


def main():
    tf.reverse = lambda x, axis: torch.flip(torch.as_tensor(np.ascontiguousarray(x)) if not isinstance(x, torch.Tensor) else x, dims=list(axis))
    ns = {"tf": tf, "np": np}
    exec(reference_function("weave_tensors"), ns)
    ns2 = {"tf": tf, "np": np, "Union": __import__("typing").Union, "Dict": dict, "Any": object}
    DS = build_class(REF, "DataSummary", ["create_statistics_index_full", "normalize"], ns2)
    ds = DS.__new__(DS)
    rng = np.random.default_rng(5600)
    out = {}
    for name, (K, T, D, H, W) in {"a": (3, 4, 2, 5, 6), "b": (2, 3, 1, 7, 9)}.items():
        Lx, Ly, Lz = 2900.0, 2900.0, 80.0
        xs = (np.arange(W) + 0.5) * (Lx / W)
        ys = (np.arange(H) + 0.5) * (Ly / H)
        zs = (np.arange(D) + 0.5) * (Lz / D)
        zg, yg, xg = np.meshgrid(zs, ys, xs, indexing="ij")
        permx = np.exp(rng.normal(np.log(3.0), 0.5, (K, D, H, W))).clip(0.26, 24.0).astype(np.float32)
        time = np.linspace(0.0, 365.0, T).astype(np.float32).reshape(T, 1)
        data = {"permx": permx, "time": time, "x": xg[None].astype(np.float32), "y": yg[None].astype(np.float32), "z": zg[None].astype(np.float32)}
        woven = ns["weave_tensors"](tensor_list=list(data.values()), target_trailing_shape=permx.shape[1:], flatten_first_axes=True,
                                    merge_consecutive_singleton_dims=True)
        woven = woven.numpy() if isinstance(woven, torch.Tensor) else np.asarray(woven)
        # statistics rows in the woven channel order [z, y, x, t, k]: [min, max, mean, std]
        stats = np.asarray([[zs.min(), zs.max(), zs.mean(), zs.std() + 1.0], [ys.min(), ys.max(), ys.mean(), ys.std()],
                            [xs.min(), xs.max(), xs.mean(), xs.std()], [0.0, 365.0, 182.5, 100.0], [0.26, 24.0, 3.0, 1.5]], np.float32)
        if D == 1:
            stats[0, 1] = stats[0, 0] + 1.0          # a single layer: keep max > min
        ds.statistics = torch.as_tensor(stats)
        cfgn = {"normalization_limits": (-1.0, 1.0), "feature_normalization_method": "lnk-linear-scaling"}
        full = torch.tensor([[0, 1, 2, 3, 4], [0, 1, 2, 3, 4]], dtype=torch.int32)
        normed = ds.normalize(torch.as_tensor(np.ascontiguousarray(woven)), norm_config=cfgn, statistics_index=full, compute=True,
                              normalization_dimension=-1, dtype=tf.float32)
        out.update({f"{name}_permx": permx, f"{name}_time": time.reshape(-1), f"{name}_x": xg.astype(np.float32), f"{name}_y": yg.astype(np.float32),
                    f"{name}_z": zg.astype(np.float32), f"{name}_stats": stats, f"{name}_woven": np.ascontiguousarray(woven),
                    f"{name}_features": normed.numpy()})
        print(name, "woven", woven.shape, "features", tuple(normed.shape))
    np.savez_compressed(os.path.join(HERE, "reference_weave.npz"), **out)
    print("wrote reference_weave.npz")


if __name__ == "__main__":
    main()
