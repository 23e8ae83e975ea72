"""Configuration of the physics-loss path.

The *values* restate the reference's defaults (default_configurations.py:20-140, 228-266,
449-451); the structure is ours: one flat ``PhysicsSpec`` that the C ABI's ``SrmConfig`` is filled
from.  ``spec_from_reference_configs`` accepts dictionaries shaped like the reference's
``DEFAULT_RESERVOIR_CONFIG`` / ``DEFAULT_WELLS_CONFIG`` / ``DEFAULT_SCAL_CONFIG`` /
``DEFAULT_GENERAL_CONFIG`` so a reference user can hand over the dictionaries they already have.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

# default_configurations.py:449-451
CONVERSION_CONSTANTS = {"field": {"C": 0.001127, "D": 5.6145833334}}

# default_configurations.py:92-130 (only the keys the path reads)
DEFAULT_RESERVOIR = {
    "porosity": 0.2, "horizontal_anisotropy": 1.0, "vertical_anisotropy": 1.0,
    "length": 2900.0, "width": 2900.0, "thickness": 80.0, "Nx": 39, "Ny": 39, "Nz": 1,
    "initialization": {"Pi": 5000, "Pa": 1000},
}

# default_configurations.py:132-140
DEFAULT_WELLS = {"connections": [
    {"name": "P1", "i": 29, "j": 29, "k": 0, "type": "producer", "control": "ORAT", "value": 500.0,
     "minimum_bhp": 4100.0, "wellbore_radius": 0.09525, "completion_ratio": 0.5, "shutin_days": [[1000.0, 0.0]]},
    {"name": "P2", "i": 29, "j": 9, "k": 0, "type": "producer", "control": "ORAT", "value": 1000.0,
     "minimum_bhp": 4100.0, "wellbore_radius": 0.09525, "completion_ratio": 0.5, "shutin_days": [[1000.0, 0.0]]},
    {"name": "P3", "i": 9, "j": 9, "k": 0, "type": "producer", "control": "ORAT", "value": 500.0,
     "minimum_bhp": 4100.0, "wellbore_radius": 0.09525, "completion_ratio": 0.5, "shutin_days": [[1000.0, 0.0]]},
    {"name": "P4", "i": 9, "j": 29, "k": 0, "type": "producer", "control": "ORAT", "value": 1000.0,
     "minimum_bhp": 4100.0, "wellbore_radius": 0.09525, "completion_ratio": 0.5, "shutin_days": [[1000.0, 0.0]]},
    {"name": "I1", "i": 19, "j": 19, "k": 0, "type": "injector", "control": "ORAT", "value": 0.0,
     "minimum_bhp": 4100.0, "wellbore_radius": 0.09525, "completion_ratio": 0.5, "shutin_days": [[1000.0, 0.0]]},
]}

# default_configurations.py:262-266
DEFAULT_SCAL = {
    "end_points": {"kro_Somax": 0.90, "krg_Sorg": 0.80, "krg_Swmin": 0.90, "Swmin": 0.22, "Sorg": 0.2,
                   "Sgc": 0.05, "Socr": 0.2, "So_max": 0.28},
    "corey_exponents": {"nog": 3.0, "ng": 6.0, "nw": 2.0},
}

# default_configurations.py:20-90 (only the keys the path reads)
DEFAULT_GENERAL = {
    "srm_start_time": 0.0, "srm_end_time": 365.0,
    "maximum_srm_timestep": 10.0, "minimum_srm_timestep": 0.1,
    "data_normalization": {"feature_normalization_method": "lnk-linear-scaling",
                           "normalization_limits": [-1.0, 1.0]},
    "physics_mode_fraction": 1.0,
    "fluid_type": "DG",
    "default_weights": {"gas": {"dom": 1.0, "ibc": 1.0, "obc": 0.0, "ic": 0.0, "td": 0.0, "mbc": 1.0,
                                "cmbc": 0.0, "tde": 1.0}},
    "srm_units": "field",
}

LOSS_KEYS = ("dom", "ibc", "obc", "ic", "td", "mbc", "cmbc", "tde")   # default_configurations.py:63-83


@dataclass
class WellSpec:
    i: int
    j: int
    k: int
    q_target: float          # signed: producer +, injector -   (welldata_processor.py:89-97)
    pwf_min: float
    rw: float
    hc: float
    shut_start: float
    shut_stop: float
    name: str = ""

    def as_dict(self):
        return dict(i=self.i, j=self.j, k=self.k, q_target=self.q_target, pwf_min=self.pwf_min, rw=self.rw,
                    hc=self.hc, shut_start=self.shut_start, shut_stop=self.shut_stop)


def wells_from_connections(connections: Sequence[dict]) -> List[WellSpec]:
    """WellDataProcessor._rebuild_tensors / get_well_data (welldata_processor.py:37-107), host side."""
    out = []
    for w in connections:
        typ = str(w.get("type", "")).strip().lower()
        sign = 1.0 if typ == "producer" else -1.0
        ctrl = str(w.get("control", "")).strip().upper()
        val = float(w.get("value", 0.0))
        q = abs(val) if ctrl == "BHP" else sign * val        # BHP always positive (:93-97)
        shut = w.get("shutin_days", [[0.0, 0.0]])
        if shut and len(shut) == 1 and len(shut[0]) == 2:    # :58-62
            s0, s1 = float(shut[0][0]), float(shut[0][1])
        else:
            s0, s1 = 0.0, 0.0
        out.append(WellSpec(i=int(w["i"]), j=int(w["j"]), k=int(w["k"]), q_target=q,
                            pwf_min=float(w.get("minimum_bhp", 0.0)), rw=float(w.get("wellbore_radius", 0.0)),
                            hc=float(w.get("completion_ratio", 0.0)), shut_start=s0, shut_stop=s1,
                            name=str(w.get("name", ""))))
    return out


def rock_compressibility_f32(phi: float) -> float:
    """cf = 97.32e-6/(1+55.8721*phi**1.428586)  (physics_loss.py:64) in fp32."""
    p = np.float32(phi)
    return float(np.float32(97.32e-6) / (np.float32(1.0) + np.float32(55.8721) * np.power(p, np.float32(1.428586))))


def corey_krog_krgo_f32(sg: float, end_points: dict, corey: dict):
    """RelativePermeability.compute_krog_krgo (relative_permeability.py:49-75) for one scalar, fp32."""
    t = np.float32
    sg = t(sg)
    swmin, sorg, sgc, socr = t(end_points["Swmin"]), t(end_points["Sorg"]), t(end_points["Sgc"]), t(end_points["Socr"])
    so = t(1.0) - sg - swmin
    with np.errstate(invalid="ignore"):
        krog = t(end_points["kro_Somax"]) * np.power((so - sorg) / (t(1.0) - swmin - sorg), t(corey["nog"]))
        krgo = t(end_points["krg_Sorg"]) * np.power((sg - sgc) / (t(1.0) - sgc - swmin - sorg), t(corey["ng"]))
    if so <= swmin + max(sorg, socr):
        krog = t(0.0)
    if sg > t(1.0) - (swmin + sorg):
        krgo = t(end_points["krg_Swmin"])
    krog = max(min(krog, t(end_points["kro_Somax"])), t(0.0))
    krgo = max(min(krgo, t(end_points["krg_Swmin"])), t(0.0))
    return float(krog), float(krgo)


@dataclass
class PhysicsSpec:
    D: int = 1
    H: int = 39
    W: int = 39
    length: float = 2900.0
    width: float = 2900.0
    thickness: float = 80.0
    phi: float = 0.2
    kx_ky: float = 1.0
    kv_kh: float = 1.0
    C: float = CONVERSION_CONSTANTS["field"]["C"]
    Dc: float = CONVERSION_CONSTANTS["field"]["D"]
    end_points: dict = field(default_factory=lambda: copy.deepcopy(DEFAULT_SCAL["end_points"]))
    corey_exponents: dict = field(default_factory=lambda: copy.deepcopy(DEFAULT_SCAL["corey_exponents"]))
    p_min: float = 14.7
    p_max: float = 10000.0
    wells: List[WellSpec] = field(default_factory=list)
    use_blocking_factor: bool = False
    n_intervals: int = 8
    root_solver: str = "newton"           # well_rate_bhp_Subclassed.py:38 ('newton' | 'chandrupatla'), GC blocking integral
    n_root_iter: int = 20                 # well_rate_bhp_Subclassed.py:40
    use_non_iterative: bool = True        # well_rate_bhp_Subclassed.py:44; False: Newton-Raphson on the BHP (_iterative_method, :515-612)
    max_iters: int = 10                   # well_rate_bhp_Subclassed.py:41
    tol: float = 1e-6                     # well_rate_bhp_Subclassed.py:42
    tde_in_dom: bool = True
    fluid_type: str = "DG"
    # time normalisation statistics (for normalize_diff of the predicted time step)
    t_min: float = 0.0
    t_max: float = 365.0
    norm_limits: tuple = (-1.0, 1.0)

    @property
    def dx(self):
        return self.length / self.W

    @property
    def dy(self):
        return self.width / self.H

    @property
    def dz(self):
        return self.thickness / self.D

    @property
    def n_cells(self):
        return self.D * self.H * self.W

    @property
    def Sgi(self):
        return float(np.float32(1.0 - self.end_points["Swmin"]))      # physics_loss.py:65

    @property
    def cf(self):
        return rock_compressibility_f32(self.phi)

    @property
    def krg(self):
        # DG: krgo at Sgi (physics_loss.py:129; well_rate_bhp_Subclassed.py:758)
        return corey_krog_krgo_f32(1.0 - self.end_points["Swmin"], self.end_points, self.corey_exponents)[1]


def scaled_default_wells(W: int, H: int, D: int = 1, all_layers: bool = False) -> List[WellSpec]:
    """The five default connections placed at the same fractional positions of a W x H grid."""
    conns = []
    for c in DEFAULT_WELLS["connections"]:
        ks = range(D) if all_layers else (0,)
        for k in ks:
            d = dict(c)
            d["i"] = min(W - 1, int(round(c["i"] / 39.0 * W)))
            d["j"] = min(H - 1, int(round(c["j"] / 39.0 * H)))
            d["k"] = k
            conns.append(d)
    return wells_from_connections(conns)


def lattice_wells(W: int, H: int, D: int, nx: int = 8, ny: int = 4) -> List[WellSpec]:
    """BASELINE config 5: nx*ny producers on a lattice, completed in every layer."""
    conns = []
    n = 0
    for a in range(ny):
        for b in range(nx):
            i = int((b + 0.5) * W / nx)
            j = int((a + 0.5) * H / ny)
            for k in range(D):
                conns.append({"name": f"L{n}", "i": i, "j": j, "k": k, "type": "producer", "control": "ORAT",
                              "value": 500.0 if (a + b) % 2 == 0 else 1000.0, "minimum_bhp": 4100.0,
                              "wellbore_radius": 0.09525, "completion_ratio": 0.5, "shutin_days": [[1000.0, 0.0]]})
            n += 1
    return wells_from_connections(conns)


def spec_from_reference_configs(reservoir: Optional[dict] = None, wells: Optional[dict] = None,
                                scal: Optional[dict] = None, general: Optional[dict] = None,
                                use_blocking_factor: bool = False, n_intervals: int = 8) -> PhysicsSpec:
    r = {**DEFAULT_RESERVOIR, **(reservoir or {})}
    g = {**DEFAULT_GENERAL, **(general or {})}
    s = {**DEFAULT_SCAL, **(scal or {})}
    w = wells if wells is not None else DEFAULT_WELLS
    units = CONVERSION_CONSTANTS[g["srm_units"]]
    lim = g["data_normalization"]["normalization_limits"]
    return PhysicsSpec(
        D=int(r["Nz"]), H=int(r["Ny"]), W=int(r["Nx"]), length=float(r["length"]), width=float(r["width"]),
        thickness=float(r["thickness"]), phi=float(r["porosity"]), kx_ky=float(r["horizontal_anisotropy"]),
        kv_kh=float(r["vertical_anisotropy"]), C=units["C"], Dc=units["D"],
        end_points=copy.deepcopy(s["end_points"]), corey_exponents=copy.deepcopy(s["corey_exponents"]),
        wells=wells_from_connections(w["connections"]), use_blocking_factor=use_blocking_factor,
        n_intervals=n_intervals, fluid_type=str(g["fluid_type"]).upper(), t_min=float(g["srm_start_time"]),
        t_max=float(g["srm_end_time"]), norm_limits=(float(lim[0]), float(lim[1])))
