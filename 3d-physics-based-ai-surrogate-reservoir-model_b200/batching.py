"""BatchGenerator: host-side mirror of the reference's batch source (training.py:16-229) with the data set RESIDENT ON
THE DEVICE.  The reference converts the whole numpy data set to a tensor inside every
``__getitem__`` (training.py:127-137); here the flattened data set is uploaded once and a batch is one row-gather kernel
(``srm_gather_rows``) over device memory.  Same constructor arguments, ``__len__``, ``__getitem__``, ``on_epoch_end``
and the same flattening of the collapsed axes (``_maybe_flatten``, Fortran order by default: b = k + K*t).

Only ``batch_axis == 0`` (the reference's default and only use) is built.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib as L


def maybe_flatten(arr: np.ndarray, collapse_axes: Sequence[int], flatten_order: str = "F") -> np.ndarray:
    """training.py:148-229 (without the optional stratified shuffle): collapse `collapse_axes` into one axis placed at
    the position of the first of them; 'F': the first collapsed axis varies fastest, 'C': the last."""
    if not collapse_axes:
        return arr
    axes = sorted(a if a >= 0 else arr.ndim + a for a in collapse_axes)
    others = [i for i in range(arr.ndim) if i not in axes]
    first = axes[0]
    if flatten_order.upper() == "C":
        if axes != list(range(axes[0], axes[0] + len(axes))):
            raise ValueError("C-order collapse needs adjacent axes (the reference reshapes in place)")
        shape = list(arr.shape)
        prod = int(np.prod([shape[a] for a in axes]))
        new_shape = shape[:first] + [prod] + shape[axes[-1] + 1:]
        flat = np.reshape(arr, new_shape)
        return np.moveaxis(flat, first, 0) if first != 0 else flat
    if flatten_order.upper() != "F":
        raise ValueError("flatten_order must be either 'C' or 'F'")
    perm = np.transpose(arr, others + axes)
    lead = [arr.shape[i] for i in others]
    # Fortran-order reshape of the WHOLE permuted array, exactly as the reference does (training.py:196-199)
    flat = np.reshape(perm, lead + [int(np.prod([arr.shape[a] for a in axes]))], order="F")
    last = len(lead)
    return np.moveaxis(flat, last, first) if first != last else flat


class BatchGenerator:
    """BatchGenerator(pairs, batch_size, collapse_axes=(0, 1), batch_axis=0, shuffle=True, stack_labels=False)

    pairs: list of (features, labels); labels an array or a dict of arrays.  ``gen[i]`` -> (x_batch, y_batch) as CUDA
    tensors (float32): x (batch, *feature_dims); y an array, a dict of arrays, or -- stack_labels -- (n_keys, batch, ...).
    """

    def __init__(self, pairs: List[tuple], batch_size: int, collapse_axes: Optional[Sequence[int]] = (0, 1),
                 batch_axis: int = 0, shuffle: bool = True, stack_labels: bool = False, device: int = 0):
        if batch_axis != 0:
            raise NotImplementedError("BatchGenerator mirror: only batch_axis = 0 (the reference's use) is built")
        if not isinstance(pairs, list):
            raise ValueError("Input 'pairs' must be a list of feature-label tuples")
        if not torch.cuda.is_available():
            raise RuntimeError("BatchGenerator (device resident) needs a CUDA device; there is no CPU fallback")
        self.batch_size, self.shuffle, self.batch_axis, self.stack_labels = int(batch_size), bool(shuffle), 0, bool(stack_labels)
        self.collapse_axes = list(collapse_axes) if collapse_axes else []
        self.device = torch.device("cuda", device)
        self.lib = L.load_library()
        self.launches = 0
        if not pairs:
            self.N, self.is_dict, self.label_keys, self.indices = 0, False, [], np.array([])
            self.x_all, self.y_all = torch.empty(0, device=self.device), torch.empty(0, device=self.device)
            return
        self.is_dict = isinstance(pairs[0][1], dict)
        self.label_keys = list(pairs[0][1].keys()) if self.is_dict else []
        if self.is_dict:
            for _, labels in pairs[1:]:
                if not isinstance(labels, dict) or set(labels.keys()) != set(self.label_keys):
                    raise ValueError("All label dictionaries must have the same keys across pairs")
        fl = lambda a: maybe_flatten(np.asarray(a), self.collapse_axes)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)
        self.x_all = up(np.concatenate([fl(f) for f, _ in pairs], axis=0))
        if self.is_dict:
            self.y_all = {k: up(np.concatenate([fl(lb[k]) for _, lb in pairs], axis=0)) for k in self.label_keys}
            if self.stack_labels:
                shapes = [tuple(self.y_all[k].shape[1:]) for k in self.label_keys]
                if not all(s == shapes[0] for s in shapes):
                    raise ValueError("All label arrays must have the same shape after flattening when stack_labels=True")
        else:
            self.y_all = up(np.concatenate([fl(lb) for _, lb in pairs], axis=0))
        self.N = int(self.x_all.shape[0])
        self.indices = np.arange(self.N)
        if self.shuffle:
            np.random.shuffle(self.indices)           # the reference's generator (training.py:103-104): numpy's global state
        self._upload_indices()

    def _upload_indices(self):
        self._idx_dev = torch.from_numpy(self.indices.astype(np.int32)).to(self.device)

    def __len__(self) -> int:
        return int(np.ceil(self.N / self.batch_size)) if self.N else 0

    def _gather(self, src: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
        n = hi - lo
        out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=self.device)
        row_bytes = src[0].numel() * src.element_size()
        idx = self._idx_dev[lo:hi]
        for a in range(0, n, 65535):                  # the kernel's grid.y limit
            b = min(n, a + 65535)
            L.check(self.lib, self.lib.srm_gather_rows(self.device.index, src.data_ptr(), idx[a:b].data_ptr(), b - a, src.shape[0],
                                                       row_bytes, out[a:b].data_ptr(), torch.cuda.current_stream(self.device).cuda_stream),
                    "srm_gather_rows")
            self.launches += 1
        return out

    def __getitem__(self, idx: int):
        if self.N == 0:
            e = torch.empty(0, device=self.device)
            return e, e
        idx = int(idx)
        lo, hi = idx * self.batch_size, min((idx + 1) * self.batch_size, self.N)
        if lo >= hi:
            raise IndexError(idx)
        x = self._gather(self.x_all, lo, hi)
        if self.is_dict:
            y = {k: self._gather(self.y_all[k], lo, hi) for k in self.label_keys}
            if self.stack_labels:
                y = torch.stack([y[k] for k in self.label_keys], dim=0)
        else:
            y = self._gather(self.y_all, lo, hi)
        return x, y

    def on_epoch_end(self):
        if self.shuffle and self.N > 0:
            np.random.shuffle(self.indices)
            self._upload_indices()
