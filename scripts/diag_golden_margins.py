"""how closely the GPU replay reproduces the records of the translated reference program (tests/golden/ref_golden.json)"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_gpu_golden import gpu_program, _cfg, GOLDEN, h
for case in json.load(open(GOLDEN))["program"]:
    cfg = _cfg(case["cfg"])
    if "Lbox_crystal" in cfg:
        cfg["Lbox"] = cfg["Lbox_crystal"]
    lat = [[h(x) for x in row] for row in case["lattice"]] if "lattice" in case else None
    e, et = gpu_program(cfg, case["Nblock"], case["Nstep"], lat)
    we = np.array([[h(x) for x in row] for row in case["e_vpi"]]); wet = np.array([[h(x) for x in row] for row in case["et_vpi"]])
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    print(f"{case['name']:9s} blocks {len(we)} x {case['Nstep']:3d} steps  Et {rel(et[:,1], wet[:,1]):.1e}  Kt {rel(et[:,2], wet[:,2]):.1e}  Vt {rel(et[:,3], wet[:,3]):.1e}"
          f"  V {rel(e[:,3], we[:,3]):.1e}  E(mixed) {rel(e[:,1], we[:,1]):.1e}  K(mixed) {rel(e[:,2], we[:,2]):.1e}")
