"""diagnostic: GPU (MT replay) vs oracle from the perfect hcp lattice start of the C4 golden case"""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pigs_oracle import Oracle
from pathintegralgroundstate_b200 import PigsCuda
h = float.fromhex
case = [c for c in json.load(open("tests/golden/ref_golden.json"))["program"] if c["name"] == "C4"][0]
c = case["cfg"]
cfg = dict(c)
for k in ("trap", "swapping", "wf_table", "v_table", "crystal"):
    cfg[k] = bool(cfg[k])
cfg["Lbox"] = cfg["Lbox_crystal"]
R = np.array([[h(x) for x in row] for row in case["lattice"]])
Nb, Np = c["Nb"], c["Np"]
P = np.broadcast_to(R, (2 * Nb + 1,) + R.shape).copy()
xe = np.stack([P[Nb, -1]] * 2)
o = Oracle(c); o.fill_tables(); W, V = o.get_tables()
o.set_state(P, xe, 0, 0); o.sgrnd(c["seed"])
g = PigsCuda(cfg, n_chains=1, rng="mt", seed=c["seed"])
which = sys.argv[1] if len(sys.argv) > 1 else "own"
if which == "own":
    g.fill_tables("hfdb")
else:
    g.set_tables(W, V)
g.sgrnd(c["seed"], chain=0)
g.set_state(0, P, xe, 0, 0)
print("Lbox", g.geo["Lbox"], "rcut", g.geo["rcut"], "oracle Lbox", o.Lbox, "tables:", which)
# 1. UpdateAction on the lattice
rng = np.random.default_rng(1)
n = 300
ip = rng.integers(1, Np + 1, n).astype(np.int32); ib = rng.integers(0, 2 * Nb + 1, n).astype(np.int32)
xold = R[ip - 1]; xnew = xold + rng.normal(0, 0.05, (n, 3))
Rr = np.broadcast_to(R, (n,) + R.shape).copy()
dg = g.update_action(Rr, ip, ib, xnew, xold)
do = np.array([o.update_action(int(a), int(b), xn, xo_, R=R) for a, b, xn, xo_ in zip(ip, ib, xnew, xold)])
err = np.abs(dg - do) / np.maximum(np.abs(do), 1e-300)
print("update_action on the lattice: max rel err", err.max(), "at", int(err.argmax()), "ib", ib[err.argmax()], dg[err.argmax()], do[err.argmax()])
# 2. blocks of one step
keys = ("idiag_block", "try_cm", "acc_cm", "try_stag", "acc_bd", "acc_head", "acc_tail", "try_open", "acc_open", "try_close", "acc_close",
        "try_swap", "acc_swap", "try_cm_half", "acc_cm_half", "try_stag_half", "acc_bd_half", "acc_head_half", "acc_tail_half")
for blk in range(4):
    g.run_block(1); bg = g.get_block(chain=0)[0]
    bo = o.run_block(1)[0]
    diff = {k: (int(bg[k]), int(bo[k])) for k in keys if int(bg[k]) != int(bo[k])}
    Po, xo2, io, iw = o.get_state(); Pg, xg, ig, iwg = g.get_state(0)
    print("step", blk + 1, "diff", diff, "isopen", ig, io, "path drift", np.max(np.abs(Po - Pg)), "updates", list(bg["bead_updates"]), list(bo["bead_updates"]))
