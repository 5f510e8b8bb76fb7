"""Throughput exploration on the GPU box (not part of the product path)."""
import sys, time, json, itertools
sys.path.insert(0, '.')
import numpy as np
from tests.common import C2, C3, CW, synthetic_path
from pathintegralgroundstate_b200 import PigsCuda, measure_fp64_peak

def run(cfg, n_chains, T, tm, nstep=4, reps=2, seed=7):
    g = PigsCuda(cfg, n_chains=n_chains, rng="philox", seed=seed, threads_per_chain=T, table_mode=tm)
    g.fill_tables()
    rng = np.random.default_rng(0)
    P0 = synthetic_path(cfg, rng, spread=0.03)
    xe0 = np.stack([P0[cfg["Nb"], -1]] * 2)
    g.set_state_all(np.broadcast_to(P0, (n_chains,) + P0.shape).copy(), np.broadcast_to(xe0, (n_chains, 2, 3)).copy())
    g.run_block(2)
    best = 0
    for _ in range(reps):
        g.run_block(nstep)
        ms = g.last_block_ms()
        b = g.get_block()[0]
        nb = sum(b["bead_updates"])
        best = max(best, nb / (ms * 1e-3))
    out = dict(Np=cfg["Np"], chains=n_chains, T=T, tm=tm, ms=ms, upd_per_s=best, idiag=b["idiag_block"], nopen=b["n_open_chains"],
               acc_bd=b["acc_bd"] / max(1, b["try_stag"]), E=b["sumE"] / max(1, b["idiag_block"]) / cfg["Np"])
    g.close()
    return out

if __name__ == "__main__":
    print("fp64 peak TF/s", measure_fp64_peak(0), flush=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    if which == "c2":
        for n, T, tm in [(4096, 32, 2), (4096, 32, 1), (4096, 32, 0), (4096, 64, 2), (2048, 64, 2), (1184, 128, 2), (592, 256, 2), (8192, 32, 2)]:
            try: print(json.dumps(run(C2, n, T, tm)), flush=True)
            except Exception as e: print("ERR", n, T, tm, e, flush=True)
    elif which == "c3":
        for n, T, tm in [(4736, 32, 2), (4736, 32, 0), (2368, 64, 2), (1184, 128, 2), (592, 256, 2), (592, 256, 0), (296, 512, 2)]:
            try: print(json.dumps(run(C3, n, T, tm, nstep=2, reps=2)), flush=True)
            except Exception as e: print("ERR", n, T, tm, e, flush=True)
