#!/usr/bin/env python3
"""Generates tests/golden/ref_golden.json from the MACHINE-TRANSLATED REFERENCE (oracle/_ref/libpigs_ref.so, built
by `make -C oracle ref` from /root/reference/*.f90 -- see oracle/f90toc/f90toc.py).  Run in the build container,
where the reference sources live:

    python tests/golden/make_ref_golden.py

The vectors are outputs of the reference's own code (not of the hand-written oracle): leaf functions, the MT19937
stream, and complete `./vpi < vpi.in` runs (e_vpi.out / et_vpi.out records, block by block).  Doubles are stored
as C99 hex strings, so tests/test_ref_pin.py::test_oracle_matches_reference_goldens compares bit for bit, also
where /root/reference does not exist (the GPU box)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pigs_ref                                    # noqa: E402
from tests.common import C1, C2, C3, CREF, CW, CWX, CS, oracle_cfg       # noqa: E402


def main():
    G = dict(source="oracle/_ref/libpigs_ref.so = f90toc translation of /root/reference/*.f90, g++ -O2 -ffp-contract=off -fwrapv",
             leaf=[], program=[])
    r = pigs_ref.Ref(oracle_cfg(C2))
    rng = np.random.default_rng(2026)
    for x in rng.uniform(0.05, 4.5, 200):
        G["leaf"].append(dict(f="potential", args=[float(x).hex()], want=r.L.ref_potential(float(x)).hex()))
        for opt in (0, 1, 2):
            G["leaf"].append(dict(f="logpsi", args=[opt, (1.2).hex(), float(x).hex()], want=r.L.ref_logpsi(opt, 1.2, float(x)).hex()))
    r.sgrnd(1982)
    G["stream"] = dict(cfg=oracle_cfg(CW), seed=1982, grnd=[r.grnd().hex() for _ in range(1300)],
                       rangauss=[r.rangauss().hex() for _ in range(300)])
    for name, cfg, Nblock, Nstep in (("CW", CW, 4, 25), ("CWX", CWX, 5, 25), ("CS", CS, 3, 20), ("C1", C1, 2, 10), ("C2", C2, 2, 2),
                                     ("CREF", CREF, 2, 3), ("C3", C3, 2, 2),      # the shipped vpi.in; the benchmarked size
                                     ("2D", dict(CWX, dim=2, density=0.3), 3, 20),      # (swapping = F walks off unallocated arrays in the reference, Q22)
                                     ("Nlev1", dict(CWX, Nlev=1, Lstag=4), 3, 20),
                                     ("CREFlong", CREF, 3, 12), ("C2worm", dict(C2, CWorm=0.5, Nobdm=5), 2, 15),    # longer runs at N = 64
                                     ("C1aniso", dict(C1, a_ho=[1.0, 1.3, 0.8]), 2, 10), ("trap2D", dict(C1, dim=2, a_ho=[1.0, 0.7, 1.0]), 2, 10),
                                     ("NlevMax", dict(CWX, Nlev=4), 3, 15), ("Nstag0", dict(CWX, Nstag=0), 3, 10)):
        c = oracle_cfg(cfg)
        rr = pigs_ref.Ref(c, Nblock=Nblock, Nstep=Nstep)
        G["program"].append(dict(name=name, cfg=c, Nblock=Nblock, Nstep=Nstep,
                                 e_vpi=[[x.hex() for x in row] for row in rr.file("e_vpi.out").tolist()],
                                 et_vpi=[[x.hex() for x in row] for row in rr.file("et_vpi.out").tolist()]))
    # BASELINE configs[3]: hcp crystal, N = 180 in an orthorhombic box, started from config_ini.in (served from memory)
    from pathintegralgroundstate_b200.workloads import config, hcp_lattice
    c4 = dict(config("C4"), CWorm=8.0)
    c4.pop("tables")
    R, Lb = hcp_lattice(density=c4["density"])
    c4["Lbox"] = [float(x) for x in Lb]
    c = oracle_cfg(c4)
    rr = pigs_ref.Ref(c, Nblock=2, Nstep=2, lattice=(R[:c4["Np"]], Lb))
    G["program"].append(dict(name="C4", cfg=c, Nblock=2, Nstep=2, lattice=[[float(x).hex() for x in row] for row in R[:c4["Np"]]],
                             e_vpi=[[x.hex() for x in row] for row in rr.file("e_vpi.out").tolist()],
                             et_vpi=[[x.hex() for x in row] for row in rr.file("et_vpi.out").tolist()]))
    # a simple-cubic lattice in a cubic box (crystal = T): neighbours at separations of EXACTLY L/2 along every axis,
    # r = rcut to the last bit -- the minimum-image comparison and the cutoff decide on equality here
    sc = dict(C2, Np=64, crystal=True, CWorm=0.5, Nobdm=4, Nstag=2)
    Lc = (64 / sc["density"]) ** (1.0 / 3.0)
    gpts = (np.arange(4) + 0.5) * (Lc / 4) - Lc / 2
    Rsc = np.array([[x, y, z] for x in gpts for y in gpts for z in gpts])
    sc["Lbox"] = [Lc, Lc, Lc]
    c = oracle_cfg(sc)
    rr = pigs_ref.Ref(c, Nblock=2, Nstep=3, lattice=(Rsc, np.array([Lc, Lc, Lc])))
    G["program"].append(dict(name="SC64", cfg=c, Nblock=2, Nstep=3, lattice=[[float(x).hex() for x in row] for row in Rsc],
                             e_vpi=[[x.hex() for x in row] for row in rr.file("e_vpi.out").tolist()],
                             et_vpi=[[x.hex() for x in row] for row in rr.file("et_vpi.out").tolist()]))
    # the same lattice with the staging flavour of every move, and a square lattice in two dimensions
    for name, base, npts, dim_ in (("SC64sta", dict(sc, sampling="sta", Lstag=6), 4, 3),
                                   ("SQ36", dict(C2, dim=2, Np=36, density=0.3, crystal=True, CWorm=0.5, Nobdm=4, Nstag=2, Nk=8), 6, 2)):
        Lq = (base["Np"] / base["density"]) ** (1.0 / dim_)
        gq = (np.arange(npts) + 0.5) * (Lq / npts) - Lq / 2
        Rq = np.array(np.meshgrid(*([gq] * dim_), indexing="ij")).reshape(dim_, -1).T.copy()
        base["Lbox"] = [Lq] * dim_
        c = oracle_cfg(base)
        rr = pigs_ref.Ref(c, Nblock=2, Nstep=3, lattice=(Rq, np.array([Lq] * dim_)))
        G["program"].append(dict(name=name, cfg=c, Nblock=2, Nstep=3, lattice=[[float(x).hex() for x in row] for row in Rq],
                                 e_vpi=[[x.hex() for x in row] for row in rr.file("e_vpi.out").tolist()],
                                 et_vpi=[[x.hex() for x in row] for row in rr.file("et_vpi.out").tolist()]))
    out = os.path.join(ROOT, "tests", "golden", "ref_golden.json")
    json.dump(G, open(out, "w"), indent=0)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
