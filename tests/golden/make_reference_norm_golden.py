"""Golden vectors for the (de)normalisation made by the REFERENCE'S OWN code: DataSummary.nonormalize and .normalize_diff
(data_processing/data_processing_utils.py:919-960, 1065-1183) are cut out by AST and executed through the torch-backed
TensorFlow stand-in on a bare instance holding a statistics table [z, y, x, time, permx, permz][min, max, mean, std].
Feature tensor channels [z, y, x, t, k] (srm_data_processing.py:668-677): linear for rows 0..3, logarithmic for the
permeability rows 4, 5 ('lnk-linear-scaling', the default).   Output: tests/golden/reference_norm.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_torch_shim as tf          # noqa: E402
from make_reference_wells_golden import build_class          # noqa: E402

REF = "/root/reference/data_processing/data_processing_utils.py"


def main():
    ns = {"tf": tf, "np": np, "Union": __import__("typing").Union, "Dict": dict, "Any": object}
    DS = build_class(REF, "DataSummary", ["create_statistics_index_full", "nonormalize", "normalize_diff"], ns)
    ds = DS.__new__(DS)
    stats = np.asarray([[40.0, 80.0, 60.0, 10.0], [37.0, 2863.0, 1450.0, 800.0], [37.0, 2863.0, 1450.0, 800.0],
                        [0.0, 365.0, 180.0, 100.0], [0.26, 24.0, 3.0, 1.5], [0.026, 2.4, 0.3, 0.15]], np.float32)
    ds.statistics = torch.as_tensor(stats)
    cfgn = {"normalization_limits": (-1.0, 1.0), "feature_normalization_method": "lnk-linear-scaling"}
    rng = np.random.default_rng(5500)
    x = (2.0 * rng.random((3, 2, 4, 5, 5)) - 1.0).astype(np.float32)
    x[0, 0, 0, 0, :] = [-1.0, 1.0, 0.0, -1.0, 1.0]
    full = torch.tensor([[0, 1, 2, 3, 4], [0, 1, 2, 3, 4]], dtype=torch.int32)
    den = ds.nonormalize(torch.as_tensor(x), norm_config=cfgn, statistics_index=full, compute=True, nonormalization_dimension=-1, dtype=tf.float32)
    tmap = torch.tensor([[0], [3]], dtype=torch.int32)
    t_only = ds.nonormalize(torch.as_tensor(x[..., 3:4]), norm_config=cfgn, statistics_index=tmap, compute=True, nonormalization_dimension=-1, dtype=tf.float32)
    dt = (0.1 + 9.9 * rng.random((3, 1, 1, 1, 1))).astype(np.float32)
    dtn = ds.normalize_diff(torch.as_tensor(dt), norm_config=cfgn, statistics_index=tmap, compute=True, nonormalization_dimension=-1, dtype=tf.float32)
    np.savez_compressed(os.path.join(HERE, "reference_norm.npz"), stats=stats, x=x, denorm=den.numpy(), t_only=t_only.numpy(), dt=dt, dt_norm=dtn.numpy())
    print("wrote reference_norm.npz", den.shape, t_only.shape, dtn.shape)


if __name__ == "__main__":
    main()
