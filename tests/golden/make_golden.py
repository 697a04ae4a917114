"""Extract the reference's own golden end states for the explicit-RHS path into small fixtures.

Run in the authoring container (where /root/reference exists):  python tests/golden/make_golden.py
Source files: /root/reference/test/CI-ref/<eqs>/<case>/output/var_<i>_0.h5 (+ t.h5), written by the
reference's write_hdf5 (src/io/write_output.jl:939-979) and compared by its CI at atol=1e-5
(test/ci_cases.jl:57,73).  The GPU box has no /root/reference, so the arrays travel as .npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle.h5mini import read_h5  # noqa: E402

REF = "/root/reference/test/CI-ref"
CASES = {"CompEuler_theta": ("CompEuler/theta", 4), "AdvDiff_kopriva": ("AdvDiff/kopriva", 1),
         "ShallowWater_SoliWaveIsland": ("ShallowWater/SoliWaveIsland", 3)}

for name, (sub, nvar) in CASES.items():
    d = os.path.join(REF, sub, "output")
    if not os.path.isdir(d):
        print("skip", name)
        continue
    out = {}
    for i in range(1, nvar + 1):
        h = read_h5(os.path.join(d, f"var_{i}_0.h5"))
        out[f"q{i}"] = h["q"]
        out[f"qe{i}"] = h["qe"]
    t = read_h5(os.path.join(d, "t.h5"))
    for k, v in t.items():
        out["t_" + k] = v
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})
