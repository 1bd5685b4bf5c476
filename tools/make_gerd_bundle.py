#!/usr/bin/env python
"""Pack the input DATA of the reference's gerd_roseires case (cases/gerd_roseires/data/*.csv) into one JSON
bundle, flow_sim_b200/cases/data/gerd_roseires.json, so the case can be built where the reference checkout
does not exist (the GPU box).  Numbers are written with repr() and therefore round-trip exactly.
Run in the build container:  python tools/make_gerd_bundle.py [/root/reference]
"""
import json
import os
import sys

import numpy as np
import pandas as pd

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
data = os.path.join(ref, "cases", "gerd_roseires", "data")
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flow_sim_b200", "cases", "data",
                   "gerd_roseires.json")


def release_table(name):
    df = pd.read_csv(os.path.join(data, name), index_col=0)
    return {"stage": df.index.to_numpy(dtype=float).tolist(), "column": df.columns.to_numpy(dtype=float).tolist(),
            "discharge": [[None if np.isnan(v) else float(v) for v in row] for row in df.to_numpy(dtype=float)]}


def hydrograph(name):          # custom_functions.import_hydrograph: skip the units row, sort by time
    t = pd.read_csv(os.path.join(data, name), skiprows=[1]).astype(np.float64).sort_values(by="time")
    return t.to_numpy().tolist()


sec = pd.read_csv(os.path.join(data, "composite_trapezoids.csv"))
cols = ["chainage", "z_min", "file", "b_main", "m_main", "b_fp_left", "b_fp_right", "m_fp", "h_bankfull", "n_left",
        "n_main", "n_right"]
coords = pd.read_csv(os.path.join(data, "centerline_coords.csv")).dropna(axis=1, how="all").dropna()
coords = coords.astype(np.float64).sort_values(by="chainage")
bundle = {
    "source": "cve-mohd/flow-sim cases/gerd_roseires/data (composite_trapezoids, inflow_hydrograph[_small], gerd_vol_curve, "
              "roseires_spillway_releases, roseires_deep_sluice_releases, centerline_coords)",
    "sections": {c: sec[c].tolist() for c in cols},
    "inflow_hydrograph_hours": hydrograph("inflow_hydrograph.csv"),
    "inflow_hydrograph_small_hours": hydrograph("inflow_hydrograph_small.csv"),
    "gerd_vol_curve": pd.read_csv(os.path.join(data, "gerd_vol_curve.csv"), header=None).to_numpy(dtype=float).tolist(),
    "spillway_releases": release_table("roseires_spillway_releases.csv"),
    "sluice_releases": release_table("roseires_deep_sluice_releases.csv"),
    "centerline": coords.to_numpy(dtype=float).tolist(),
}
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "w") as f:
    json.dump(bundle, f, separators=(",", ":"))
print(out, os.path.getsize(out), "bytes")
