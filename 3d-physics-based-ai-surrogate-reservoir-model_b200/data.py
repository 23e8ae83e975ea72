"""Feature construction and the on-disk formats either side of it.

  weave_features   the reference's `weave_tensors` (data_processing/data_processing_utils.py:90-223, call site
                   srm_data_processing.py:363-403) + `DataSummary.normalize` (:979-1063) in one CUDA pass: the
                   (K*T, D, H, W, 5) feature tensor, channels [z, y, x, t, k] in [lo, hi], is written once in HBM
                   (srm_weave_features); nothing of shape (K, T, D, H, W, 5) ever exists on the host
  read_permx_dat / write_permx_dat
                   the keyword files of the KLE realisation generator (kle_realization_generator.py:179-229):
                   comment lines, the keyword, one value per line (n*value repeats accepted), a closing "/"
  read_kle_npy     the stacked realisations (kle_realization_generator.py:230-260: np.save / np.savez_compressed)
  positional_grids cell-centre coordinates of a regular grid in the (D, H, W) layout (srm_data_processing.py:320-361)

The simulator label files (.FUNRST / .RSM, simulation_data_process_pipeline.py:148-296) feed the validation plots
only -- nothing on the physics-loss path reads them -- and are not read here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def positional_grids(D: int, H: int, W: int, length: float, width: float, thickness: float):
    """cell-centre coordinates (z, y, x), each (D, H, W) fp32"""
    xs = (np.arange(W, dtype=np.float64) + 0.5) * (length / W)
    ys = (np.arange(H, dtype=np.float64) + 0.5) * (width / H)
    zs = (np.arange(D, dtype=np.float64) + 0.5) * (thickness / D)
    zg, yg, xg = np.meshgrid(zs, ys, xs, indexing="ij")
    return zg.astype(np.float32), yg.astype(np.float32), xg.astype(np.float32)


def weave_features(permx: torch.Tensor, time: torch.Tensor, x: torch.Tensor, y: torch.Tensor, z: torch.Tensor,
                   stats, limits: Tuple[float, float] = (-1.0, 1.0)) -> torch.Tensor:
    """permx (K, D, H, W), time (T,), x / y / z (D, H, W): contiguous fp32 CUDA tensors on one device.
    stats: (5, >=2) rows [z, y, x, t, k] with (min, max) first -- DataSummary.statistics in the woven channel order.
    Returns the normalised features (K*T, D, H, W, 5), realisation-major (sample b = k*T + t)."""
    if not torch.cuda.is_available():
        raise RuntimeError("weave_features needs a CUDA device; there is no CPU fallback")
    lib = L.load_library()
    dev = permx.device
    for t, nm in ((permx, "permx"), (time, "time"), (x, "x"), (y, "y"), (z, "z")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == dev and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"{nm}: need a contiguous fp32 CUDA tensor on {dev}")
    K = permx.shape[0]
    grid = tuple(permx.shape[1:])
    cells = int(np.prod(grid))
    if x.numel() != cells or y.numel() != cells or z.numel() != cells:
        raise ValueError("coordinate grids do not match permx's trailing shape")
    T = time.numel()
    st = torch.as_tensor(np.asarray(stats, dtype=np.float32)[:5, :2].copy(), device=dev).contiguous()
    if st.shape != (5, 2):
        raise ValueError("stats: need five rows [z, y, x, t, k] of (min, max, ...)")
    out = torch.empty((K * T,) + grid + (5,), dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    L.check(lib, lib.srm_weave_features(dev.index or 0, K, T, cells, p(permx), p(time), p(x), p(y), p(z), p(st), float(limits[0]),
                                        float(limits[1]), p(out), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
            "srm_weave_features")
    return out


def write_permx_dat(path: str, field: np.ndarray, keyword: str = "PERMX", comments: Sequence[str] = (), prefix: str = "--"):
    """kle_realization_generator.py:179-229: comments, keyword, one value per line in C order, '/'"""
    with open(path, "w") as f:
        for c in comments:
            f.write(f"{prefix} {c}\n")
        f.write(f"{keyword}\n")
        for v in np.asarray(field).reshape(-1):
            f.write(f"{v}\n")
        f.write("/\n")


def read_permx_dat(path: str, shape: Optional[Sequence[int]] = None, keyword: Optional[str] = None, prefix: str = "--") -> np.ndarray:
    """values of one keyword block as fp32 (reshaped to `shape` if given); accepts several values per line and the
    simulator's n*value repeat counts"""
    vals = []
    seen = keyword is None
    with open(path) as f:
        for line in f:
            s = line.split(prefix, 1)[0].strip()
            if not s:
                continue
            if not seen:
                seen = s.split()[0].upper() == keyword.upper()
                continue
            done = False
            for tok in s.split():
                if tok == "/":
                    done = True
                    break
                if tok.endswith("/"):
                    tok, done = tok[:-1], True
                if "*" in tok:
                    n, v = tok.split("*", 1)
                    vals.extend([float(v)] * int(n))
                else:
                    try:
                        vals.append(float(tok))
                    except ValueError:
                        if vals:
                            raise ValueError(f"{path}: unexpected token {tok!r} inside the data block")
                        # the keyword line of a file read without naming the keyword
                if done:
                    break
            if done:
                break
    if not seen:
        raise ValueError(f"{path}: keyword {keyword!r} not found")
    a = np.asarray(vals, dtype=np.float32)
    if shape is not None:
        if a.size != int(np.prod(shape)):
            raise ValueError(f"{path}: {a.size} values, expected {int(np.prod(shape))} for shape {tuple(shape)}")
        a = a.reshape(tuple(shape))
    return a


def read_kle_npy(path: str, key: Optional[str] = None) -> np.ndarray:
    """(K, Nz, Ny, Nx) realisations from the generator's .npy / .npz dump"""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    a = np.load(path)
    if isinstance(a, np.lib.npyio.NpzFile):
        k = key or a.files[0]
        a = a[k]
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 3:
        a = a[None]
    if a.ndim != 4:
        raise ValueError(f"{path}: expected (K, Nz, Ny, Nx) or (Nz, Ny, Nx), got {a.shape}")
    return np.ascontiguousarray(a)
