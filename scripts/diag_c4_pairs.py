"""diagnostic 2: which partners make GPU and oracle UpdateAction differ on the perfect hcp lattice"""
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
o = Oracle(c); o.fill_tables(); W, V = o.get_tables()
g = PigsCuda(cfg, n_chains=1, rng="mt", seed=c["seed"])
g.set_tables(W, V)
L = np.array(g.geo["Lbox"]); rcut = g.geo["rcut"]
print("L", L.tolist(), "rcut", repr(rcut), "L/2", (L / 2).tolist(), "inside box:", np.abs(R).max(axis=0).tolist())
rng = np.random.default_rng(1)
n = 300
ip = rng.integers(1, Np + 1, n).astype(np.int32); ib = rng.integers(0, 2 * Nb + 1, n).astype(np.int32)
xold = R[ip - 1]; xnew = xold + rng.normal(0, 0.05, (n, 3))
Rr = np.broadcast_to(R, (n,) + R.shape).copy()
dg = g.update_action(Rr, ip, ib, xnew, xold)
do = np.array([o.update_action(int(a), int(b), xn, xo_, R=R) for a, b, xn, xo_ in zip(ip, ib, xnew, xold)])
err = np.abs(dg - do)
bad = np.argsort(-err)[:6]
print("abs err sorted:", err[bad], "ib", ib[bad])
print("how many evaluations differ by more than 1e-9:", int((err > 1e-9).sum()), "of", n, " by slice class: even", int(((err > 1e-9) & (ib % 2 == 0) & (ib > 0) & (ib < 2 * Nb)).sum()),
      "odd", int(((err > 1e-9) & (ib % 2 == 1)).sum()), "end", int(((err > 1e-9) & ((ib == 0) | (ib == 2 * Nb))).sum()))
i = bad[0]
def mimg(d):
    d = d.copy()
    for k in range(3):
        d[:, k] = np.where(d[:, k] > L[k] / 2, d[:, k] - L[k], d[:, k])
        d[:, k] = np.where(d[:, k] < -L[k] / 2, d[:, k] + L[k], d[:, k])
    return d
for nm, x in (("old", xold[i]), ("new", xnew[i])):
    d = mimg(x[None, :] - R)
    r = np.sqrt((d * d).sum(axis=1))
    near = np.flatnonzero(np.abs(r - rcut) < 1e-6)
    print(nm, "partners within 1e-6 of rcut:", [(int(j), float(r[j] - rcut), d[j].tolist()) for j in near if j != ip[i] - 1])
# remove the suspicious partners one at a time (move them far inside the cutoff exclusion: put them at distance > rcut clearly)
d = mimg(xold[i][None, :] - R); r = np.sqrt((d * d).sum(axis=1))
near = [int(j) for j in np.flatnonzero(np.abs(r - rcut) < 1e-6) if j != ip[i] - 1]
for j in near:
    R2 = R.copy(); R2[j] += np.array([0.3, 0.2, 0.0])          # clearly off the sphere
    a = g.update_action(R2[None], ip[i:i + 1], ib[i:i + 1], xnew[i:i + 1], xold[i:i + 1])[0]
    b = o.update_action(int(ip[i]), int(ib[i]), xnew[i], xold[i], R=R2)
    print("partner", j, "moved off the sphere: gpu", a, "oracle", b, "diff", a - b)
print("original: gpu", dg[i], "oracle", do[i], "diff", dg[i] - do[i], "wS*V(rcut)-ish:", 2 * c["dt"] / 3 * V[-2], V[-3:], "table len", len(V))
