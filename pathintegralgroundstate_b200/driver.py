"""`program vpi` on top of libpigs_cuda: the reference driver's outer layers
(vpi.f90:76-653) -- input, geometry, tables, initial path, block loop,
normalisation, statistics, output files, checkpoint -- with the step loop
(vpi.f90:297-475) replaced by `pigs_run_block`.

This is the tested stand-in for the Fortran driver + `fortran/pigs_cuda_mod.f90`
(no Fortran compiler exists in this image).  It writes the reference's files:
e_vpi.out, et_vpi.out ('(5g20.10e3)', vpi.f90:517-518), gr_vpi.out, sk_vpi.out,
nr_vpi.out ('(20g20.10e3)', sample_mod.f90:803-864), jastrow.out, potential.out
(vpi_mod.f90:96-139), fort.99 (vpi.f90:590-592), checkpoint.dat
(vpi_mod.f90:263-309) and rand_state (random_mod.f90:125-158, gfortran
unformatted records), and can resume from them (vpi_mod.f90:162-185).

The compute backend is any object with PigsCuda's methods; the GPU tests run the
same driver over the CPU oracle through an adapter and compare the files.

With n_chains > 1 every sum the reference keeps per block is the sum over
chains (the normalisers are linear in them): block averages become averages
over chains x diagonal steps.
"""
from __future__ import annotations

import math
import os
import struct
import time

import numpy as np

from .host import derive_geometry, make_table, mcmillan_logpsi, aziz_hfdb, aziz_hfdhe2, kn_ball


# ------------------------------------------------------------------ Fortran edit descriptors
def fortran_f(x: float, w: int, d: int) -> str:
    """Fw.d (gfortran spells non-finite values NaN / Infinity, or Inf when the field is narrower than 8)"""
    if math.isnan(x):
        return "NaN".rjust(w)
    if math.isinf(x):
        t = ("Infinity" if w >= 8 else "Inf") if x > 0 else ("-Infinity" if w >= 9 else "-Inf")
        return t.rjust(w) if len(t) <= w else "*" * w
    s = f"{x:.{d}f}"
    if d == 0:
        s += "."                       # Fortran always prints the decimal point
    if s.startswith("0.") and len(s) > w:
        s = s[1:]
    elif s.startswith("-0.") and len(s) > w:
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def _scaled(a: float, ex: int) -> float:
    """a / 10**ex without overflow or underflow of the power at the ends of the double range"""
    if ex > 300:
        return a / 1e300 / 10.0 ** (ex - 300)
    if ex < -300:
        return a * 1e300 / 10.0 ** (ex + 300)
    return a / 10.0 ** ex


def fortran_e(x: float, w: int, d: int, e: int) -> str:
    """Ew.dEe with the default scale factor 0: 0.dddddE+eee"""
    if x == 0.0 or not math.isfinite(x):
        if math.isnan(x):
            return "NaN".rjust(w)
        if math.isinf(x):
            return ("Infinity" if x > 0 else "-Infinity").rjust(w)
        mant, ex = "0." + "0" * d, 0
        sgn = "-" if math.copysign(1.0, x) < 0 else ""
    else:
        sgn = "-" if x < 0 else ""
        a = abs(x)
        ex = int(math.floor(math.log10(a))) + 1
        ms = f"{_scaled(a, ex):.{d}f}"
        if ms.startswith("1."):            # rounding carried into the next decade
            ex += 1
            ms = f"{_scaled(a, ex):.{d}f}"
        mant = ms
    s = f"{sgn}{mant}E{'+' if ex >= 0 else '-'}{abs(ex):0{e}d}"
    if len(s) > w and s.lstrip("-").startswith("0."):
        s = s.replace("0.", ".", 1)
    return s.rjust(w) if len(s) <= w else "*" * w


def fortran_g(x: float, w: int = 20, d: int = 10, e: int = 3) -> str:
    """Gw.dEe (Fortran 2003 10.6.4.1.2): F editing with e+2 trailing blanks when the
    magnitude fits d significant digits, E editing otherwise."""
    n = e + 2
    if not math.isfinite(x):
        return fortran_e(x, w, d, e)
    a = abs(x)
    if a == 0.0:
        return fortran_f(x, w - n, d - 1) + " " * n
    if a < 0.1 - 0.5 * 10.0 ** (-d - 1) or a >= 10.0 ** d - 0.5:
        return fortran_e(x, w, d, e)
    # find k with 10^(k-1) - 0.5*10^(k-d-1) <= a < 10^k - 0.5*10^(k-d)
    k = int(math.floor(math.log10(a))) + 1
    if a >= 10.0 ** k - 0.5 * 10.0 ** (k - d):
        k += 1
    elif a < 10.0 ** (k - 1) - 0.5 * 10.0 ** (k - d - 1):
        k -= 1
    k = min(max(k, 0), d)
    return fortran_f(x, w - n, d - k) + " " * n


def g_line(values, w=20, d=10, e=3) -> str:
    return "".join(fortran_g(float(v), w, d, e) for v in values) + "\n"


def var(nitem: int, s: float, s2: float) -> float:
    """Var (sample_mod.f90:921-932): sqrt((<x^2>-<x>^2)/n), NaN when rounding makes the radicand negative"""
    r = (s2 - s * s) / float(nitem)
    return math.sqrt(r) if r >= 0 else float("nan")


# ------------------------------------------------------------------ checkpoint files
def write_checkpoint(path, trap, Path, xend, isopen, iworm):
    """CheckPoint (vpi_mod.f90:263-309): list-directed text, particle-major, bead-minor"""
    S, Np, dim = Path.shape
    with open(path, "w") as f:
        f.write(" .True.\n" if trap else " .False.\n")
        f.write(" .True.\n" if isopen else " .False.\n")
        f.write(f" {int(iworm):11d}\n")
        for ip in range(Np):
            for ib in range(S):
                f.write(" " + " ".join(f"{Path[ib, ip, k]:24.16E}" for k in range(dim)) + "\n")
        f.write("\n\n")
        for j in range(2):
            f.write(" " + " ".join(f"{xend[j, k]:24.16E}" for k in range(dim)) + "\n")


def read_checkpoint(path, dim, Np, Nb):
    """init, resume branch (vpi_mod.f90:162-185)"""
    toks = open(path).read().split()
    tf = lambda t: t.strip(".").lower().startswith("t")
    trap, isopen, iworm = tf(toks[0]), tf(toks[1]), int(toks[2])
    S = 2 * Nb + 1
    v = np.array([float(t.replace("D", "E")) for t in toks[3:3 + Np * S * dim + 2 * dim]])
    Path = v[:Np * S * dim].reshape(Np, S, dim).transpose(1, 0, 2).copy()
    xend = v[Np * S * dim:].reshape(2, dim).copy()
    return trap, isopen, iworm, Path, xend


def append_rand_state(path, mt, mti):
    """mtsavef(fname,'u') (random_mod.f90:125-158): two gfortran unformatted sequential records, APPENDED"""
    with open(path, "ab") as f:
        f.write(struct.pack("<iii", 4, int(mti), 4))
        body = np.asarray(mt, dtype="<u4").tobytes()
        f.write(struct.pack("<i", len(body)) + body + struct.pack("<i", len(body)))


def read_rand_state(path):
    """mtgetf(fname,'u') (random_mod.f90:162-191): reads the FIRST (oldest) pair of records (Q10)"""
    with open(path, "rb") as f:
        n0, mti, n1 = struct.unpack("<iii", f.read(12))
        assert n0 == 4 and n1 == 4
        (n,) = struct.unpack("<i", f.read(4))
        mt = np.frombuffer(f.read(n), dtype="<u4").copy()
    return mt, mti


# ------------------------------------------------------------------ the driver
# ------------------------------------------------------------------ stdout of `program vpi` (vpi.f90:161-194, 552-586, 620-634)
# list-directed `print *` starts every record with one blank (gfortran) and writes a default integer in 12 columns;
# formats 101-105 are the reference's: 101 (x,a,x,f7.2,x,a)  102 (a,x,G16.8e2,x,a,x,G16.8e2)  103 (x,a,x,i5)
# 104 (x,a,x,G13.6e2)  105 (x,a,x,3G13.6e2).  Written from the Fortran 2003 rules; no gfortran exists here to diff against.
def _ld(text="", ival=None):
    return " " + text + ("" if ival is None else f"{ival:12d}")


def _f101(text, val=None, tail=None):
    return " " + text + ("" if val is None else " " + fortran_f(val, 7, 2) + " " + tail)


def _f102(text, a, b):
    return text + " " + fortran_g(a, 16, 8, 2) + " +/- " + fortran_g(b, 16, 8, 2)


def banner_lines(cfg, geo, Nblock, Nstep, n_chains, schedule=None):
    sta = str(cfg["sampling"]).strip().lower().startswith("sta")
    out = [_ld(""), _ld("=============================================================="),
           _ld("                      VPI Monte Carlo                         "),
           _ld("=============================================================="), _ld(""), _ld(" "),
           _ld("# The Monte Carlo sampling will be performed using " + ("STAGING" if sta else "BISECTION")), _ld("  algorithm"),
           _ld("# The Monte Carlo sampling will use swap updates" if cfg.get("swapping") else
               "# The Monte Carlo sampling will not use swap updates"),
           _ld(" "), _ld("# Simulation parameters:"), _ld(""),
           f" {'  > Dimensions          :'} {geo['dim']:5d}", f" {'  > Number of particles :'} {geo['Np']:5d}"]
    if geo["trap"]:
        out.append(" " + "  > Trapping length     :" + " " + "".join(fortran_g(a, 13, 6, 2) for a in geo["a_ho"][:geo["dim"]]))
    else:
        out.append(" " + "  > Density             :" + " " + fortran_g(geo["density"], 13, 6, 2))
        out.append(" " + "  > Size of the box     :" + " " + "".join(fortran_g(a, 13, 6, 2) for a in geo["Lbox"][:geo["dim"]]))
    out += [f" {'  > Number of beads     :'} {int(cfg['Nb']):5d}", " " + "  > Time step           :" + " " + fortran_g(float(cfg["dt"]), 13, 6, 2),
            f" {'  > Number of blocks    :'} {Nblock:5d}", f" {'  > MC steps per block  :'} {Nstep:5d}"]
    if n_chains > 1:        # not in the reference: the replicas this build runs at once
        out.append(f" {'  > Markov chains (GPU) :'} {n_chains:5d}")
    out.append(_ld(""))
    return out


def block_report_lines(iblock, m, bvar, b, Np, Nstep, n_chains, dt):
    f32 = lambda x: float(np.float32(x))
    pct = lambda a, t: (f32(100 * a) / t) if t else float("nan")                     # 100*real(acc)/try   (try is real(8))
    pct2 = lambda a, t: (100.0 * f32(a) / f32(t)) if t else float("nan")             # 100.d0*real(acc)/real(try)
    out = [_ld("-----------------------------------------------------------"), _ld("BLOCK NUMBER :", iblock), _ld(" "),
           _ld("# Block results:"), _ld(" ")]
    for lab, i in (("  > <E>  =", 0), ("  > <Ec> =", 1), ("  > <Ep> =", 2)):
        out.append(_f102(lab, m[i] / Np, bvar[i] / Np))
    out.append(_ld(" "))
    for lab, i in (("  > <Et> =", 3), ("  > <Kt> =", 4), ("  > <Vt> =", 5)):
        out.append(_f102(lab, m[i] / Np, bvar[i] / Np))
    out += [_ld(""), _ld("# Acceptance of diagonal movements:"), _ld(" "),
            _f101("> CM movements      =", pct(b["acc_cm"], b["try_cm"]), "%"),
            _f101("> Staging movements =", pct(b["acc_bd"], b["try_stag"]), "%"),
            _f101("> Head movements    =", pct(b["acc_head"], b["try_stag"]), "%"),
            _f101("> Tail movements    =", pct(b["acc_tail"], b["try_stag"]), "%"),
            _ld(" "), _ld("# Acceptance of off-diagonal movements:"), _ld(" "),
            _f101("> CM movements      =", pct(b["acc_cm_half"], b["try_cm_half"]), "%"),
            _f101("> Staging movements =", pct(b["acc_bd_half"], b["try_stag_half"]), "%"),
            _f101("> Head movements    =", pct(b["acc_head_half"], b["try_stag_half"]), "%"),
            _f101("> Tail movements    =", pct(b["acc_tail_half"], b["try_stag_half"]), "%"),
            _ld(" "), _ld("# Acceptance open/close updates:"), _ld(" "),
            _f101("> Diagonal conf.    =", pct2(b["idiag_block"], Nstep * n_chains), "%"),
            _f101("> Open acc          =", pct2(b["acc_open"], b["try_open"]), "%"),
            _f101("> Close acc         =", pct2(b["acc_close"], b["try_close"]), "%"),
            _f101("> Swap acc          =", pct2(b["acc_swap"], b["try_swap"]), "%"),
            _f101(" "), _f101("# Time per block    =", dt, "seconds"),
            f" # GPU throughput    = {sum(b['bead_updates']) / dt:.4g} bead-updates/s"]
    return out


def final_lines(A, V, Np):
    out = [_ld("=============================================================="), _ld("FINAL RESULTS:"), _ld(""), _ld("# Final averages:"), _ld("")]
    for lab, i in (("  > <E>  =", 0), ("  > <Ec> =", 1), ("  > <Ep> =", 2)):
        out.append(_f102(lab, A[i] / Np, V[i] / Np))
    out.append(_ld(""))
    for lab, i in (("  > <Et> =", 3), ("  > <Kt> =", 4), ("  > <Vt> =", 5)):
        out.append(_f102(lab, A[i] / Np, V[i] / Np))
    out += [_ld(""), _ld("=============================================================="), _ld("")]
    return out


class VpiDriver:
    def __init__(self, cfg: dict, backend, workdir: str = ".", potential: str = "hfdb", quiet: bool = False):
        self.cfg, self.be, self.wd, self.quiet = dict(cfg), backend, workdir, quiet
        self.geo = derive_geometry(cfg)
        self.n_chains = getattr(backend, "n_chains", 1)
        self.potential = potential
        os.makedirs(workdir, exist_ok=True)
        self.out = []

    def say(self, *a):
        line = " ".join(str(x) for x in a)
        self.out.append(line)
        if not self.quiet:
            print(line)

    def _p(self, name):
        return os.path.join(self.wd, name)

    # JastrowTable / PotentialTable incl. jastrow.out / potential.out (vpi_mod.f90:84-145)
    def tables(self):
        c, g = self.cfg, self.geo
        Nmax = g["Nmax"]
        pot = dict(hfdb=aziz_hfdb, hfdhe2=aziz_hfdhe2, zero=lambda r: np.zeros_like(r))[self.potential]
        W = make_table(lambda r: mcmillan_logpsi(r, float(c["Rm"])), g["rcut"], Nmax) if self.potential != "zero" \
            else np.zeros(Nmax + 2)
        V = make_table(pot, g["rcut"], Nmax)
        r = np.arange(Nmax) * g["dr"]
        with np.errstate(all="ignore"):
            with open(self._p("jastrow.out"), "w") as f:
                for i in range(Nmax):
                    f.write(g_line([r[i], math.exp(W[i + 1]) if np.isfinite(W[i + 1]) else 0.0, W[i + 1]]))
            with open(self._p("potential.out"), "w") as f:
                for i in range(Nmax):
                    f.write(g_line([r[i], V[i + 1]]))
        self.be.set_tables(W, V)
        return W, V

    # init (vpi_mod.f90:149-259)
    def init_state(self):
        c, g, be = self.cfg, self.geo, self.be
        dim, Np, Nb = g["dim"], g["Np"], int(c["Nb"])
        S = 2 * Nb + 1
        self.full_resume = bool(c.get("resume")) and hasattr(be, "load_checkpoint") and \
            os.path.exists(self._p("checkpoint_chains.bin")) and os.path.exists(self._p("checkpoint_driver.bin"))
        if self.full_resume:
            # the extended checkpoint of an earlier run: every chain (library) and every accumulator (run())
            be.load_checkpoint(self._p("checkpoint_chains.bin"))
            return
        if c.get("resume"):
            trap, isopen, iworm, Path, xend = read_checkpoint(self._p("checkpoint.dat"), dim, Np, Nb)
            mt, mti = read_rand_state(self._p("rand_state"))
            for ch in range(self.n_chains):
                be.set_state(ch, Path, xend, int(isopen), iworm)
            if hasattr(be, "set_mt"):
                be.set_mt(0, mt, mti)
            return
        seed = int(c.get("seed", 1982))
        for ch in range(self.n_chains):
            be.sgrnd(seed + ch, chain=ch)
            if c.get("crystal"):
                R = np.loadtxt(self._p("config_ini.in"), skiprows=3)[:Np, :dim]        # vpi_mod.f90:220-228
            else:
                u = be.grnd(Np * dim, chain=ch).reshape(Np, dim)                        # ip-major, k-minor draws
                if g["trap"]:
                    R = 2.0 * np.asarray(g["a_ho"][:dim]) * (u - 0.5)                   # vpi_mod.f90:208-214
                else:
                    R = np.asarray(g["Lbox"][:dim]) * (u - 0.5)                         # vpi_mod.f90:232-236
            Path = np.broadcast_to(R, (S, Np, dim)).copy()
            xend = np.stack([Path[Nb, Np - 1], Path[Nb, Np - 1]])                        # vpi_mod.f90:250-254
            be.set_state(ch, Path, xend, 0, 0)

    def run(self, Nblock=None, Nstep=None):
        c, g, be = self.cfg, self.geo, self.be
        Nblock = int(c["Nblock"] if Nblock is None else Nblock)
        Nstep = int(c["Nstep"] if Nstep is None else Nstep)
        dim, Np, Nb = g["dim"], g["Np"], int(c["Nb"])
        Nbin, Nk, Npw = int(c["Nbin"]), int(c.get("Nk", 0)), int(c.get("Npw", 0))
        Nobdm, CWorm, trap = int(c.get("Nobdm", 0)), float(c.get("CWorm", 0.0)), g["trap"]
        density, rbin = g["density"], g["rbin"]
        self.tables()
        self.init_state()
        for ln in banner_lines(c, g, Nblock, Nstep, self.n_chains):      # vpi.f90:161-194
            self.say(ln)
        Av = np.zeros(6)
        Av2 = np.zeros(6)
        AvGr, AvGr2 = np.zeros(Nbin), np.zeros(Nbin)
        AvSk, AvSk2 = np.zeros((Nk, dim)), np.zeros((Nk, dim))
        AvNr, AvNr2 = np.zeros((Nbin, Npw + 1)), np.zeros((Nbin, Npw + 1))
        nrho = np.zeros((Nbin, Npw + 1))
        idiag_aux, obdm_bl, diag_bl = 0, 0, 0
        k_n = kn_ball(dim)
        rr = (np.arange(1, Nbin + 1, dtype=np.float64) - 0.5) * rbin
        nid = density * k_n * ((rr + 0.5 * rbin) ** dim - (rr - 0.5 * rbin) ** dim)
        iblock0 = 0
        arrs = [Av, Av2, AvGr, AvGr2, AvSk, AvSk2, AvNr, AvNr2, nrho]

        def driver_state(save):
            # checkpoint_driver.bin: "PIGSDRV1", uint64 count, doubles [iblock, idiag_aux, obdm_bl, diag_bl, Av, Av2, AvGr,
            # AvGr2, AvSk, AvSk2, AvNr, AvNr2, nrho] -- the same file the compiled driver (csrc/vpi_main.cpp) writes
            nonlocal iblock0, idiag_aux, obdm_bl, diag_bl
            n = 4 + sum(a.size for a in arrs)
            if save:
                v = np.concatenate([[iblock0, idiag_aux, obdm_bl, diag_bl]] + [a.ravel() for a in arrs]).astype("<f8")
                with open(self._p("checkpoint_driver.bin"), "wb") as f:
                    f.write(b"PIGSDRV1" + np.uint64(v.size).tobytes() + v.tobytes())
            else:
                raw = open(self._p("checkpoint_driver.bin"), "rb").read()
                if raw[:8] != b"PIGSDRV1" or int(np.frombuffer(raw[8:16], "<u8")[0]) != n:
                    raise ValueError("checkpoint_driver.bin does not belong to this configuration")
                v = np.frombuffer(raw[16:16 + 8 * n], "<f8")
                iblock0, idiag_aux, obdm_bl, diag_bl = int(v[0]), int(v[1]), int(v[2]), int(v[3])
                o = 4
                for a in arrs:
                    a[...] = v[o:o + a.size].reshape(a.shape)
                    o += a.size

        if getattr(self, "full_resume", False):
            driver_state(False)
        mode = "a" if getattr(self, "full_resume", False) else "w"
        fe = open(self._p("e_vpi.out"), mode)
        fet = open(self._p("et_vpi.out"), mode)
        self.blocks = []
        every = int(c.get("checkpoint_every", 1))
        for iblock in range(iblock0 + 1, Nblock + 1):
            t0 = time.perf_counter()
            be.run_block(Nstep)
            b, gr, Sk, nr = be.get_block()
            nd = int(b["idiag_block"])
            nrho += nr
            idiag_aux += nd
            if nd != 0:                                                     # vpi.f90:477-520
                m = np.array([b["sumE"], b["sumK"], b["sumV"], b["sumEt"], b["sumKt"], b["sumVt"]]) / float(nd)
                m2 = np.array([b["sumE2"], b["sumK2"], b["sumV2"], b["sumEt2"], b["sumKt2"], b["sumVt2"]]) / float(nd)
                bvar = [var(nd, m[i], m2[i]) for i in range(6)]
                diag_bl += 1
                Av += m
                Av2 += m * m
                if not trap:
                    ngr = int(b["ngr"])
                    grn = gr / (nid * (float(Np) * float(ngr)))            # NormalizeGr (sample_mod.f90:656-679)
                    Skn = Sk / (float(Np) * float(ngr))                    # NormalizeSk (:683-702)
                    AvGr += grn
                    AvGr2 += grn * grn
                    AvSk += Skn
                    AvSk2 += Skn * Skn
                fe.write(g_line([np.float32(iblock), m[0] / Np, m[1] / Np, m[2] / Np]))
                fet.write(g_line([np.float32(iblock), m[3] / Np, m[4] / Np, m[5] / Np]))
            else:
                m, bvar = np.full(6, np.nan), [np.nan] * 6
            # OBDM block (vpi.f90:522-539): closes once >= Nstep diagonal configurations (per chain) accumulated
            if idiag_aux // (Nstep * self.n_chains) >= 1:
                obdm_bl += 1
                if not trap and CWorm > 0 and Nobdm > 0:
                    nrn = nrho / (CWorm * nid * float(idiag_aux) * float(Nobdm))[:, None]     # NormalizeNr (:706-732)
                    AvNr += nrn
                    AvNr2 += nrn * nrn
                idiag_aux = 0
                nrho[:] = 0.0
            self.checkpoint()
            if hasattr(be, "save_checkpoint") and every > 0 and (iblock % every == 0 or iblock == Nblock):
                be.save_checkpoint(self._p("checkpoint_chains.bin"))
                iblock0 = iblock
                driver_state(True)
                fe.flush()
                fet.flush()
            dt = time.perf_counter() - t0
            self.blocks.append(dict(b, means=m, time=dt))
            for ln in block_report_lines(iblock, m, bvar, b, Np, Nstep, self.n_chains, dt):      # vpi.f90:552-586
                self.say(ln)
        fe.close()
        fet.close()
        # fort.99: permutation histogram (vpi.f90:590-592), summed over chains
        hist = np.zeros(Np, dtype=np.int64)
        for ch in range(self.n_chains):
            hist += be.get_perm(ch)[2]
        with open(self._p("fort.99"), "w") as f:
            for ip in range(Np):
                f.write(f" {ip + 1:11d} {int(hist[ip]):11d}\n")
        # finals (vpi.f90:606-642)
        res = {}
        if diag_bl:
            A, A2 = Av / diag_bl, Av2 / diag_bl
            V = [var(diag_bl, A[i], A2[i]) for i in range(6)]
            for ln in final_lines(A, V, Np):                                   # vpi.f90:620-634
                self.say(ln)
            res = dict(E=A[0] / Np, K=A[1] / Np, V=A[2] / Np, Et=A[3] / Np, Kt=A[4] / Np, Vt=A[5] / Np,
                       errE=V[0] / Np, errEt=V[3] / Np)
        if not trap and diag_bl:
            with open(self._p("gr_vpi.out"), "w") as f:                     # NormAvGr (sample_mod.f90:794-816)
                a, a2 = AvGr / diag_bl, AvGr2 / diag_bl
                for j in range(Nbin):
                    f.write(g_line([rr[j], a[j], var(diag_bl, a[j], a2[j])]))
            with open(self._p("sk_vpi.out"), "w") as f:                     # NormAvSk (:820-842)
                a, a2 = AvSk / diag_bl, AvSk2 / diag_bl
                qb = [2.0 * g["pi"] / g["Lbox"][k] for k in range(dim)]
                for j in range(Nk):
                    row = []
                    for k in range(dim):
                        row += [(j + 1) * qb[k], a[j, k], var(diag_bl, a[j, k], a2[j, k])]
                    f.write(g_line(row))
            if obdm_bl:
                with open(self._p("nr_vpi.out"), "w") as f:                 # NormAvNr (:846-870)
                    a, a2 = AvNr / obdm_bl, AvNr2 / obdm_bl
                    for j in range(Nbin):
                        row = [rr[j]]
                        for mm in range(Npw + 1):
                            row += [a[j, mm], var(obdm_bl, a[j, mm], a2[j, mm])]
                        f.write(g_line(row))
        res["diag_bl"], res["obdm_bl"] = diag_bl, obdm_bl
        return res

    def checkpoint(self):
        """CheckPoint of chain 0 every block (vpi.f90:541-545) in the reference's single-chain format"""
        Path, xend, isopen, iworm = self.be.get_state(0)
        write_checkpoint(self._p("checkpoint.dat"), self.geo["trap"], Path, xend, isopen, iworm)
        if hasattr(self.be, "get_mt"):
            mt, mti = self.be.get_mt(0)
            append_rand_state(self._p("rand_state"), mt, mti)


def main(argv=None):
    """`python -m pathintegralgroundstate_b200.driver < vpi.in` -- the reference's command line"""
    import argparse
    import sys
    from .host import PigsCuda
    from .vpi_in import read_vpi_in
    ap = argparse.ArgumentParser(description="PIGS/VPI Monte Carlo on B200: reads the reference's vpi.in from stdin")
    ap.add_argument("--workdir", default=".")
    ap.add_argument("--chains", type=int, default=None)
    ap.add_argument("--rng", default=None, choices=["philox", "mt"])
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=None, help="shard the chains over this many GPUs (inside the C ABI)")
    ap.add_argument("--potential", default="hfdb", choices=["hfdb", "hfdhe2", "zero"])
    a = ap.parse_args(argv)
    cfg = read_vpi_in(sys.stdin.read())
    if cfg.get("crystal"):                       # Np, Lbox, density come from config_ini.in (vpi.f90:101-107)
        from .host import read_config_ini
        cfg["Np"], cfg["Lbox"], cfg["density"], _ = read_config_ini(os.path.join(a.workdir, "config_ini.in"), int(cfg["dim"]))
    cu = cfg.get("cuda", {})
    n = a.chains or int(cu.get("n_chains", 1))
    rng = a.rng or str(cu.get("rng", "mt" if n == 1 else "philox"))
    sim = PigsCuda(cfg, n_chains=n, rng=rng, device=a.device, threads_per_chain=int(cu.get("threads_per_chain", 0)),
                   table_mode=int(cu.get("table_mode", -1)), schedule=int(cu.get("schedule", -1)),
                   gpus=a.gpus or int(cu.get("gpus", 1)))
    if "checkpoint_every" in cu:
        cfg["checkpoint_every"] = int(cu["checkpoint_every"])
    VpiDriver(cfg, sim, workdir=a.workdir, potential=a.potential).run()


if __name__ == "__main__":
    main()
