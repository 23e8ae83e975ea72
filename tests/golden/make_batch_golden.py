"""Golden vectors for the BatchGenerator mirror, made by the REFERENCE'S OWN code: `BatchGenerator._maybe_flatten` and
the index arithmetic of `__getitem__` are cut out of /root/reference/training.py by AST (the module itself imports
TensorFlow and cannot be imported here) and run on seeded arrays.  Output: tests/golden/batch_golden.npz.

    python tests/golden/make_batch_golden.py        (in the build container, where /root/reference exists)
"""
import ast
import os
import textwrap

import numpy as np

REF = "/root/reference/training.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_maybe_flatten():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BatchGenerator")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "_maybe_flatten")
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    ns = {"np": np}
    exec(code, ns)
    return ns["_maybe_flatten"]


class _Self:
    def __init__(self, collapse_axes):
        self.collapse_axes = list(collapse_axes)


def main():
    mf = reference_maybe_flatten()
    rng = np.random.default_rng(4100)
    out = {}
    cases = {"kt_dhw5": ((3, 4, 2, 3, 5, 5), (0, 1)), "kt_only": ((5, 7), (0, 1)), "three": ((2, 3, 4, 6), (0, 1, 2)),
             "inner": ((4, 2, 3, 5), (1, 2))}
    for name, (shape, axes) in cases.items():
        a = rng.standard_normal(shape).astype(np.float32)
        out[f"{name}_in"] = a
        out[f"{name}_axes"] = np.asarray(axes)
        out[f"{name}_F"] = mf(_Self(axes), a, flatten_order="F")
        if name != "inner" or True:
            out[f"{name}_C"] = mf(_Self(axes), a, flatten_order="C")
    # the batching arithmetic of __getitem__ (training.py:121-125) on a shuffled index vector
    N, bs = 23, 5
    np.random.seed(4101)
    ind = np.arange(N)
    np.random.shuffle(ind)
    out["idx_N"], out["idx_bs"], out["idx_perm"] = np.asarray(N), np.asarray(bs), ind
    out["idx_batches"] = np.asarray([[ind[min(i * bs + j, N - 1)] if i * bs + j < min((i + 1) * bs, N) else -1 for j in range(bs)]
                                     for i in range(int(np.ceil(N / bs)))])
    np.savez_compressed(os.path.join(HERE, "batch_golden.npz"), **out)
    print("wrote batch_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
