#!/usr/bin/env python3
"""One pinned CPU process of the reference arm of bench.py (TEST / BASELINE INFRASTRUCTURE, oracle/).

Runs the reference program `./vpi < vpi.in` -- oracle/_ref/libpigs_ref.so, the reference's Fortran sources machine-
translated to C++ and compiled with g++ -O2 (no Fortran compiler exists in the image) -- for one independent chain,
as BASELINE.md section 3 prescribes: one process per host core, pinned, own seed.  Falls back to the hand-written
oracle port when the translated library is absent.

The program keeps no state between runs, so the K timed steps are isolated by running it twice from the same seed,
with W and with W + K Monte-Carlo steps: the difference is exactly the K steps that follow the W warm-up steps.

    python -m oracle.ref_worker <core> <seed> <W> <K> <workload> [port|native]
prints one JSON line {"updates": ..., "seconds": ..., "kind": ...}.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    core, seed, W, K, workload = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    impl = sys.argv[6] if len(sys.argv) > 6 else "ref"
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    import numpy as np
    from pathintegralgroundstate_b200.workloads import config, lattice_sites
    from pathintegralgroundstate_b200.host import derive_geometry
    cfg = config(workload)
    cfg = {k: (int(v) if isinstance(v, bool) else v) for k, v in cfg.items() if k != "tables"}
    cfg["seed"] = seed
    geo = derive_geometry(cfg)
    rng = np.random.default_rng(seed)
    R = lattice_sites(cfg) + rng.uniform(-0.05, 0.05, size=(cfg["Np"], 3))
    L = np.asarray(geo["Lbox"])
    R = (R + L / 2) % L - L / 2
    from oracle import pigs_ref
    if impl == "ref" and os.path.exists(pigs_ref.LIB):
        pigs_ref.REFERENCE = "/nonexistent"          # never rebuild inside a timed worker
        def run(nstep):
            t0 = time.perf_counter()
            r = pigs_ref.Ref(cfg, Nblock=1, Nstep=nstep, lattice=(R, L))
            return r.bead_updates(), time.perf_counter() - t0
        u0 = 0
        uA, tA = run(W) if W > 0 else (0, 0.0)
        uB, tB = run(W + K)
        # the counter is cumulative over the process: run A counted uA, run B counted uB - uA
        out = dict(updates=(uB - uA) - uA, seconds=tB - tA, kind="reference (machine-translated Fortran -> C++, g++ -O2)")
    else:
        from oracle.pigs_oracle import Oracle
        o = Oracle(cfg, native=(impl == "native"))
        o.fill_tables()
        P = np.broadcast_to(R, (2 * cfg["Nb"] + 1,) + R.shape).copy()
        o.set_state(P, np.stack([R[-1], R[-1]]), 0, 0)
        o.sgrnd(seed)
        if W > 0:
            o.run_block(W)
        t0 = time.perf_counter()
        b, _, _, _ = o.run_block(K)
        out = dict(updates=int(sum(b["bead_updates"])), seconds=time.perf_counter() - t0,
                   kind="port (hand-written C++ restatement, g++ " + ("-O3 -march=x86-64-v3" if impl == "native" else "-O2") + ")")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
