"""Golden vectors for the PVT spline and the relative permeabilities made by the REFERENCE'S OWN code, executed through
the torch-backed TensorFlow stand-in of this directory (see tf_torch_shim.py; TensorFlow is not installable here):

  * polyhm_splines.py is executed whole (module source, unmodified); PolyharmonicSplineInterpolationLayer is instantiated
    on the PVT table and called on seeded pressures, orders 1 and 2;
  * PVTLayer (PVT_Layer_Subclassed.py:22-216) is cut out by AST and called: clamp, value stack, derivative by the tape
    (torch autograd stands in for TF's: the derivative is recorded but only loosely comparable, its rounding noise
    depends on the framework's accumulation order);
  * RelativePermeability (relative_permeability.py) is executed whole and compute_krog_krgo called on seeded
    saturations.

The stand-in's matmul accumulates the inner index sequentially in fp32 (multiply, then add): the order the oracle pins.
Output: tests/golden/reference_pvt_relperm.npz
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import tf_torch_shim as tf          # noqa: E402
import srm_oracle as O              # noqa: E402

REF = "/root/reference"


def exec_module(path, extra=None):
    sys.modules["tensorflow"] = tf
    ns = {"__name__": "reference_module"}
    ns.update(extra or {})
    exec(compile(open(path).read(), path, "exec"), ns)
    return ns


def cut_class(path, name):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == name)
    return textwrap.dedent(ast.get_source_segment(src, node))


def main():
    out = {}
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    rng = np.random.default_rng(5200)
    sp = exec_module(os.path.join(REF, "polyhm_splines.py"))
    Layer = sp["PolyharmonicSplineInterpolationLayer"]
    knots = np.asarray(cols["Pre"], np.float32)
    p = np.concatenate([4100.0 + 900.0 * rng.random(300), knots[(knots > 3000) & (knots < 6000)] + rng.normal(0, 0.5, ((knots > 3000) & (knots < 6000)).sum()),
                        [14.7, 10000.0, 1000.0, 2500.0, 7000.0]]).astype(np.float32)
    q = torch.as_tensor(p).reshape(1, -1, 1, 1)
    out["p"] = p
    for order in (1, 2):
        tab = O.build_spline_table(cols, O.GC_PROPS, order=order, lam=0.001)
        for pi, prop in enumerate(O.GC_PROPS):
            vals = np.asarray(tab.f[pi], np.float32)
            lay = Layer(train_points=knots, train_values=vals, order=order, regularization_weight=0.001, name=f"{prop}_spline")
            y_full = lay(q)                                                     # the layer's own solve + evaluation
            w_ref, v_ref = Layer._solve_interpolation(lay.train_points, lay.train_values, order, 0.001)
            w_o = torch.as_tensor(tab.w[pi], dtype=torch.float32).reshape(1, -1, 1)
            v_o = torch.as_tensor(tab.v[pi], dtype=torch.float32).reshape(1, 2, 1)
            y_wv = Layer._apply_interpolation(q.reshape(1, -1, 1), lay.train_points, w_o, v_o, order)   # the oracle's (w, v) as data
            out[f"o{order}_{prop}_full"] = y_full.reshape(-1).detach().numpy()
            out[f"o{order}_{prop}_wv"] = y_wv.reshape(-1).detach().numpy()
            out[f"o{order}_{prop}_w"] = w_ref.reshape(-1).numpy()
            out[f"o{order}_{prop}_v"] = v_ref.reshape(-1).numpy()
    # PVTLayer.call (dry gas): clamp + stack + tape derivative
    lookup = {"pre": knots}
    tab1 = O.build_spline_table(cols, O.DG_PROPS, order=1, lam=0.001)
    for pi, prop in enumerate(["invBg", "invug"]):
        lookup[prop] = np.asarray(tab1.f[pi], np.float32)
    ns = {"tf": tf, "np": np, "PolyharmonicSplineInterpolationLayer": Layer}
    exec(cut_class(os.path.join(REF, "PVT_Layer_Subclassed.py"), "PVTLayer"), ns)
    cfg = types.SimpleNamespace(lookup=lambda k: lookup[k])
    layer = ns["PVTLayer"](fluid_type="DG", fitting_method="spline", spline_config=cfg, spline_order=1, regularization_weight=0.001)
    pin = torch.as_tensor(np.concatenate([p, [5.0, 12000.0]]).astype(np.float32)).reshape(1, -1, 1, 1)
    res = layer(pin)
    out["pvt_p"] = pin.reshape(-1).numpy()
    out["pvt_layer_out"] = res.detach().numpy()                                # [2, n_prop, B, m, 1, 1]
    # relative permeability
    rp = exec_module(os.path.join(REF, "relative_permeability.py"), {"np": np})
    RP = rp["RelativePermeability"]
    ocfg = O.OracleConfig()
    ep = dict(Swmin=ocfg.Swmin, Sorg=ocfg.Sorg, Sgc=ocfg.Sgc, Socr=ocfg.Socr, So_max=ocfg.So_max, kro_Somax=ocfg.kro_Somax, krg_Sorg=ocfg.krg_Sorg, krg_Swmin=ocfg.krg_Swmin)
    model = RP(end_points=ep, corey_exponents=dict(nog=ocfg.nog, ng=ocfg.ng), dtype=tf.float32)
    sg = np.concatenate([rng.random(400), [0.0, 0.05, 0.2, 0.58, 0.5799999, 0.78, 0.7800001, 1.0, 0.38, 0.3800001]]).astype(np.float32)
    krog, krgo = model.compute_krog_krgo(torch.as_tensor(sg))
    out["sg"], out["krog"], out["krgo"] = sg, krog.numpy(), krgo.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_pvt_relperm.npz"), **out)
    print("wrote reference_pvt_relperm.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
