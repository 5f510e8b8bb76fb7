"""Freezes outputs of the CPU oracle as regression vectors (tests/golden/oracle_golden.json).
The reference is Fortran 90 and cannot be built or imported in this image, so these
are NOT reference outputs: they guard the oracle against accidental change.
    python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pigs_oracle import Oracle  # noqa: E402
from tests.common import CW, oracle_cfg, synthetic_path  # noqa: E402

cfg = dict(CW)
rng = np.random.default_rng(2026)
o = Oracle(oracle_cfg(cfg))
o.fill_tables()
P = synthetic_path(cfg, rng, spread=0.03)
xe = np.stack([P[cfg["Nb"], -1]] * 2)
o.set_state(P, xe, 0, 0)
seed = 4242
o.sgrnd(seed)
uas = []
for _ in range(12):
    ip = int(rng.integers(1, cfg["Np"] + 1))
    ib = int(rng.integers(0, 2 * cfg["Nb"] + 1))
    xold = P[ib, ip - 1]
    xnew = xold + rng.normal(0, 0.1, 3)
    uas.append(dict(ip=ip, ib=ib, xnew=xnew.tolist(), xold=xold.tolist(), dS=o.update_action(ip, ib, xnew, xold)))
nstep = 8
b, gr, Sk, nr = o.run_block(nstep)
ints = {k: int(v) for k, v in b.items() if k.startswith(("acc_", "try_", "idiag", "ngr"))}
json.dump(dict(cfg=cfg, seed=seed, path=P.tolist(), xend=xe.tolist(), update_action=uas, nstep=nstep, block_int=ints,
               bead_updates=list(b["bead_updates"]), sumE=b["sumE"], sumEt=b["sumEt"], gr=gr.tolist()),
          open(os.path.join(os.path.dirname(__file__), "oracle_golden.json"), "w"))
print("wrote golden:", ints)
