"""Golden vectors for the simulator label readers made by the REFERENCE'S OWN parsers: the module
data_processing/simulation_data_process_pipeline.py is plain numpy Python and imports here, so `parse_continuous_file`
(:247-292) and `parse_tabular_file_from_string` (:148-245) are run as they are on two small synthetic files written in
the formats they read -- a formatted restart file (.FUNRST: quoted keyword headers, free-format numbers, several report
steps, a logical block and an integer header in between) and a run summary (.RSM: tab-separated segmented tables with
SUMMARY banners, three header lines per table, a compound "WOPR of the well at 15 15 1" column, a cell that is not a
number and an empty cell).  The page-break "1" lines of a real .RSM are left out: a numeric line with no header above
it sends the reference parser into an endless loop (:186-187 `continue` without advancing); the reader here skips them.

Output: tests/golden/reference_simfiles.npz (the two texts and every parsed array)
        python tests/golden/make_reference_simfiles_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/data_processing")
import simulation_data_process_pipeline as ref          # noqa: E402

RSM_SPEC = [["TIME"], ["WOPR", "15 15 1"], ["WOPR", "20  20 1"], "WGPR", "WWPR", "WBHP", "FPR"]
RST_KEYS = ["PRESSURE", "SGAS", "SOIL", "SWAT"]


def restart_text(rng, D, H, W, steps):
    n = D * H * W
    out = []
    for s in range(steps):
        out.append(f" 'SEQNUM  '           1 'INTE'")
        out.append(f"           {s + 1}")
        out.append(f" 'INTEHEAD'           6 'INTE'")
        out.append("  -1620679  201901  1  0  0  0")
        out.append(f" 'LOGIHEAD'           4 'LOGI'")
        out.append("  F  T  F  F")
        for key, lo, hi in (("PRESSURE", 4100.0, 5000.0), ("SGAS", 0.5, 0.78), ("SOIL", 0.0, 0.28)):
            v = rng.uniform(lo, hi, n)
            out.append(f" '{key:8s}'        {n:4d} 'REAL'")
            for i in range(0, n, 4):
                out.append("  " + "  ".join(f"{x:.8E}" for x in v[i:i + 4]))
        if s == 1:
            out.append("")          # an empty line closes a block too
    return "\n".join(out) + "\n"


def rsm_text(rng, rows):
    t = np.linspace(0.0, 365.0, rows)
    # the well-name and cell-index header lines start in column 0: the reference strips leading tabs from every line, so
    # header cells behind empty leading columns would slide to the left (table 2 below pins exactly that)
    tab1 = ["\tSUMMARY OF RUN CASE_A", "\tWOPR\tWOPR\tWGPR\tTIME\tYEARS\tFPR", "\tSTB/DAY\tSTB/DAY\tMSCF/DAY\tDAYS\tYEARS\tPSIA",
            "\tP1\tP2\tP1", "\t15 15 1\t20 20  1\t15 15 1"]
    for i in range(rows):
        cells = [f"{rng.uniform(0, 900):.3f}", f"{rng.uniform(0, 900):.3f}", f"{rng.uniform(0, 5000):.3f}", f"{t[i]:.4f}", f"{t[i] / 365.25:.6f}",
                 f"{rng.uniform(4100, 5000):.3f}"]
        if i == 2:
            cells[0] = "*******"          # overflowed field: not a number
        if i == 3:
            cells[2] = ""                 # empty cell: skipped
        tab1.append("\t" + "\t".join(cells))
    tab2 = ["\tSUMMARY OF RUN CASE_A", "\tTIME\tWWPR\tWBHP\tWBHP", "\tDAYS\tSTB/DAY\tPSIA\tPSIA", "\t\tP1\tP1\tP2"]
    for i in range(rows):
        tab2.append("\t" + "\t".join([f"{t[i]:.4f}", f"{rng.uniform(0, 50):.4f}", f"{rng.uniform(4100, 4900):.2f}", f"{rng.uniform(4100, 4900):.2f}"]))
    tab3 = ["\tSUMMARY OF RUN CASE_A", "\tDATE\tNEWTON\tMSUMLINS", "\t\tITERS\tITERS"]          # nothing requested in here
    for i in range(rows):
        tab3.append("\t" + "\t".join([f"{i + 1}", f"{rng.integers(1, 9)}", f"{rng.integers(10, 99)}"]))
    return "\n".join(tab1 + ["", ""] + tab2 + [""] + tab3) + "\n"


def main():
    rng = np.random.default_rng(5700)
    out = {}
    rst = restart_text(rng, 2, 3, 5, 3)
    got = ref.parse_continuous_file(rst, RST_KEYS)
    out["restart_text"] = np.asarray(rst)
    for k in RST_KEYS:
        out[f"restart_{k}_n"] = np.asarray(len(got[k]))
        for i, a in enumerate(got[k]):
            out[f"restart_{k}_{i}"] = a
    print("restart:", {k: len(v) for k, v in got.items()})
    rsm = rsm_text(rng, 6)
    tab = ref.parse_tabular_file_from_string(rsm, RSM_SPEC)
    out["rsm_text"] = np.asarray(rsm)
    for k, v in tab.items():
        if isinstance(v, dict):
            for s, a in v.items():
                out[f"rsm_{k}|{s}"] = np.asarray([]) if a is None else a
                out[f"rsm_{k}|{s}_none"] = np.asarray(a is None)
        else:
            out[f"rsm_{k}"] = np.asarray([]) if v is None else v
            out[f"rsm_{k}_none"] = np.asarray(v is None)
    print("rsm:", {k: (None if v is None else ({s: (None if a is None else a.shape) for s, a in v.items()} if isinstance(v, dict) else v.shape))
                   for k, v in tab.items()})
    np.savez_compressed(os.path.join(HERE, "reference_simfiles.npz"), **out)
    print("wrote reference_simfiles.npz")


if __name__ == "__main__":
    main()
