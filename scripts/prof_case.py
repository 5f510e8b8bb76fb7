"""One short production (Philox) run of the sweep kernel, for ncu (not part of the product path).
usage: prof_case.py <c2|c3> <n_chains> <threads_per_chain> <table_mode> <nstep> [warm_blocks]"""
import sys
sys.path.insert(0, '.')
import numpy as np
from tests.common import C2, C3, CWX, synthetic_path
from pathintegralgroundstate_b200 import PigsCuda

cfg = dict(c2=C2, c3=C3, cw=CWX)[sys.argv[1]]
n, T, tm, nstep = (int(x) for x in sys.argv[2:6])
warm = int(sys.argv[6]) if len(sys.argv) > 6 else 1
g = PigsCuda(cfg, n_chains=n, rng="philox", seed=7, threads_per_chain=T, table_mode=tm)
g.fill_tables()
rng = np.random.default_rng(0)
P0 = synthetic_path(cfg, rng, spread=0.03)
xe0 = np.stack([P0[cfg["Nb"], -1]] * 2)
g.set_state_all(np.broadcast_to(P0, (n,) + P0.shape).copy(), np.broadcast_to(xe0, (n, 2, 3)).copy())
for _ in range(warm):
    g.run_block(nstep)
g.run_block(nstep)
b = g.get_block()[0]
print("plan", g.launch_plan(), "ms", g.last_block_ms(), "upd/s", sum(b["bead_updates"]) / (g.last_block_ms() * 1e-3), "launches", g.launch_count())
