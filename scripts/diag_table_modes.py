import json, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from tests.test_gpu_golden import gpu_program, _cfg, GOLDEN, h
G = json.load(open(GOLDEN))["program"]
for name in ("SC64", "C4", "CWX"):
    case = [c for c in G if c["name"] == name][0]
    cfg = _cfg(case["cfg"])
    if "Lbox_crystal" in cfg: cfg["Lbox"] = cfg["Lbox_crystal"]
    lat = [[h(x) for x in row] for row in case["lattice"]] if "lattice" in case else None
    wet = np.array([[h(x) for x in row] for row in case["et_vpi"]])
    for tm in (0, 1, 2):
        try:
            e, et = gpu_program(cfg, case["Nblock"], case["Nstep"], lat, table_mode=tm)
            ok = et.shape == wet.shape and np.allclose(et[:, 1:], wet[:, 1:], rtol=1e-9, atol=1e-9 * np.abs(wet[:, 1:]).max())
            print(name, "table_mode", tm, "OK" if ok else f"DIFFERS shapes {et.shape} {wet.shape}")
        except Exception as ex:
            print(name, "table_mode", tm, "error", ex)
