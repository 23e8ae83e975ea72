#!/usr/bin/env python
"""Export the reference's PVT table to a travelling fixture.

Reads ``/root/reference/pvt_data.df`` (pandas pickle, 37 rows x 10 float32 columns; loaded by
``default_configurations.py:545-560`` in the reference) and writes it bit-exactly as
``pvt_table.npz`` next to this script and into the package's ``data/`` directory.

This script only runs in the build container (the reference is not present on the GPU box);
its output is committed.  Usage:  python tests/golden/make_pvt_table.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "3d-physics-based-ai-surrogate-reservoir-model_b200")


def main(ref="/root/reference/pvt_data.df"):
    import pandas as pd

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = pd.read_pickle(ref)
    cols = list(df.columns)
    table = np.stack([df[c].to_numpy() for c in cols], axis=0)
    assert table.dtype == np.float32 and table.shape == (10, 37), (table.dtype, table.shape)
    out = dict(columns=np.array(cols), table=table)
    for d in (HERE, os.path.join(PKG, "data")):
        os.makedirs(d, exist_ok=True)
        np.savez(os.path.join(d, "pvt_table.npz"), **out)
    print("wrote pvt_table.npz:", cols, table.shape)


if __name__ == "__main__":
    main(*sys.argv[1:])
