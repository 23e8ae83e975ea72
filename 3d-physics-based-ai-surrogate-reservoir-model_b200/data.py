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

  read_restart_keywords / read_rsm_columns
                   the simulator's label files (simulation_data_process_pipeline.py:148-296): formatted restart /
                   init files (.FUNRST, .FINIT: quoted keyword headers followed by free-format numbers, one block
                   per report step) and the tab-separated run summary (.RSM: segmented tables whose column titles
                   span several header lines).  They feed the validation plots only -- nothing on the physics-loss
                   path reads them; host-side text parsing, pinned to the reference's own parsers by goldens
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

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


# ---- simulator label files (validation data) ---------------------------------------------------------------------
def _text_of(src: str) -> str:
    """a path to read, or the file's text itself (anything holding a newline is taken as text)"""
    if "\n" not in src and os.path.exists(src):
        with open(src) as f:
            return f.read()
    return src


def read_restart_keywords(src: str, keys: Sequence[str], dtype=np.float32) -> Dict[str, List[np.ndarray]]:
    """Formatted (ASCII) restart / init file -> {keyword: [one 1-D array per block, in file order]}.

    parse_continuous_file, simulation_data_process_pipeline.py:247-292: a line that starts with a quote opens the block of
    the keyword between its first two quotes (` 'PRESSURE'  841 'REAL'`); the block's numbers run until the next header or
    an empty line; a line inside a wanted block that holds a non-number is dropped whole; blocks without numbers
    are not reported.  Reshape with `restart_to_grid`."""
    want = set(keys)
    out: Dict[str, List[np.ndarray]] = {k: [] for k in keys}
    name, vals = None, []

    def close():
        if name in want and vals:
            out[name].append(np.asarray(vals, dtype=dtype))

    for raw in _text_of(src).splitlines():
        line = raw.strip()
        if line[:1] == "'":
            close()
            q = line.split("'")
            name, vals = (q[1].strip() if len(q) > 1 else None), []
        elif not line:
            close()
            name, vals = None, []
        elif name in want:
            try:
                vals.extend([float(t) for t in line.split()])
            except ValueError:
                pass
    close()
    return out


def restart_to_grid(blocks: Sequence[np.ndarray], D: int, H: int, W: int) -> np.ndarray:
    """report-step blocks of D*H*W cell values (simulator order: x fastest, then y, then layers) -> (steps, D, H, W)"""
    if not blocks:
        return np.zeros((0, D, H, W), np.float32)
    for b in blocks:
        if b.size != D * H * W:
            raise ValueError(f"block of {b.size} values, expected {D * H * W} for a {W}x{H}x{D} grid")
    return np.stack([np.asarray(b).reshape(D, H, W) for b in blocks])


def _numeric_row(line: str, threshold: float = 0.7) -> bool:
    """is_mostly_numbers (:93-99): at least `threshold` of the non-empty tab-separated cells parse as floats"""
    cells = [c.strip() for c in line.split("\t")]
    cells = [c for c in cells if c]
    if not cells:
        return False
    n = 0
    for c in cells:
        try:
            float(c)
            n += 1
        except ValueError:
            pass
    return n / len(cells) >= threshold


def _column_titles(header: Sequence[str]) -> List[str]:
    """merge_header_lines (:101-118): the first header line fixes the column count; the cells of the following lines are
    appended to their column's title; titles are compared with runs of white space collapsed"""
    cols = [c.strip() for c in header[0].split("\t")]
    for line in header[1:]:
        cells = [c.strip() for c in line.split("\t")]
        for i, c in enumerate(cells[:len(cols)]):
            if c:
                cols[i] += " " + c
    return [" ".join(c.split()) for c in cols]


def _column_spec(spec) -> Dict[str, Union[List[str], Dict[str, List[str]]]]:
    """convert_target_spec (:120-146): "WGPR" -> {WGPR: [WGPR]}; ["WOPR", "15 15 1"] -> {WOPR: {"15 15 1": [WOPR, 15 15 1]}}"""
    if isinstance(spec, dict):
        return spec
    d: Dict[str, Union[List[str], Dict[str, List[str]]]] = {}
    for item in spec:
        if isinstance(item, str):
            d[item] = [item]
        elif len(item) < 2:
            d[item[0]] = [item[0]]
        else:
            sub = " ".join(item[1:]).strip()
            if not isinstance(d.get(item[0]), dict):
                d[item[0]] = {}
            d[item[0]][sub] = list(item)
    return d


def read_rsm_columns(src: str, spec, dtype=np.float32):
    """Run-summary (.RSM) tables -> {name: array | None} (compound entries: {name: {qualifier: array | None}}).

    parse_tabular_file_from_string (:148-245).  The file is a sequence of tables separated by empty lines: `SUMMARY`
    banner lines are skipped, the lines up to the first mostly-numeric one form the header, the numeric lines that
    follow are the rows (tab-separated).  A requested column is the FIRST one whose merged title contains every phrase
    of its spec; values of the same column in later tables are appended; a cell that is not a number gives NaN, an
    empty or missing cell is skipped; a name with no value at all yields None."""
    want = _column_spec(spec)
    got: Dict[str, Union[list, Dict[str, list]]] = {k: ({s: [] for s in v} if isinstance(v, dict) else []) for k, v in want.items()}
    lines = [ln.lstrip("\t").rstrip() for ln in _text_of(src).split("\n")]
    n, i = len(lines), 0
    banner = lambda ln: ln.strip().upper().startswith("SUMMARY")

    def first_with(titles, phrases):
        ph = [" ".join(p.split()) for p in phrases]
        return next((c for c, t in enumerate(titles) if all(p in t for p in ph)), None)

    def put(bucket, col, cells):
        if col < len(cells) and cells[col]:
            try:
                bucket.append(float(cells[col]))
            except ValueError:
                bucket.append(np.nan)

    while i < n:
        while i < n and (not lines[i].strip() or banner(lines[i])):
            i += 1
        header = []
        while i < n and lines[i].strip() and not _numeric_row(lines[i]):
            if not banner(lines[i]):
                header.append(lines[i].strip())
            i += 1
        if i >= n and not header:
            break
        if not header:
            if i < n and _numeric_row(lines[i]):
                i += 1          # rows without a header above them belong to no table
            continue
        titles = _column_titles(header)
        cols: Dict[str, Union[int, Dict[str, int]]] = {}
        for k, v in want.items():
            if isinstance(v, dict):
                cols[k] = {s: c for s, c in ((s, first_with(titles, ph)) for s, ph in v.items()) if c is not None}
            else:
                c = first_with(titles, v)
                if c is not None:
                    cols[k] = c
        if not cols or all(isinstance(c, dict) and not c for c in cols.values()):
            while i < n and lines[i].strip():
                i += 1
            continue
        while i < n and not lines[i].strip():
            i += 1
        while i < n and lines[i].strip() and _numeric_row(lines[i]):
            cells = [c.strip() for c in lines[i].split("\t")]
            for k, c in cols.items():
                if isinstance(c, dict):
                    for s, cc in c.items():
                        put(got[k][s], cc, cells)
                else:
                    put(got[k], c, cells)
            i += 1
    fin = lambda v: np.asarray(v, dtype=dtype) if v else None
    return {k: ({s: fin(x) for s, x in v.items()} if isinstance(v, dict) else fin(v)) for k, v in got.items()}
