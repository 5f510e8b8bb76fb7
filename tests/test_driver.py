"""The driver layer (vpi.f90 outer loop on top of the C ABI): Fortran edit
descriptors, checkpoint/rand_state formats, and the whole driver run over the
CPU oracle (CPU test) and over libpigs_cuda (GPU test) with the files compared."""
import os
import subprocess

import numpy as np
import pytest

from pathintegralgroundstate_b200.driver import (fortran_g, fortran_e, fortran_f, g_line, var, write_checkpoint,
                                                 read_checkpoint, append_rand_state, read_rand_state, VpiDriver)
from pathintegralgroundstate_b200.vpi_in import format_vpi_in, read_vpi_in
from pathintegralgroundstate_b200 import derive_geometry
from tests.common import C1, CWX, CW
from tests.oracle_backend import OracleBackend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VPI = os.path.join(ROOT, "pathintegralgroundstate_b200", "vpi_cuda")


def _vpi(args, stdin, cwd=None):
    """run the compiled driver (csrc/vpi_main.cpp, built by csrc/Makefile / __graft_entry__.build)"""
    if not os.path.exists(VPI):       # normally built by csrc/Makefile; host-only C++, so a direct g++ call is enough
        pkg = os.path.dirname(VPI)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(pkg, "csrc", "vpi_main.cpp"),
                               "-o", VPI, "-L" + pkg, "-l:libpigs_cuda.so", "-Wl,-rpath,$ORIGIN"])
    r = subprocess.run([VPI] + list(args), input=stdin, capture_output=True, text=True, cwd=cwd, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_fortran_g_editing():
    # Gw.dEe: F editing + e+2 trailing blanks inside [0.1, 10^d), E editing outside (F2003 10.6.4.1.2)
    assert fortran_g(1.0) == "    1.000000000     "
    assert fortran_g(0.5) == "   0.5000000000     "
    assert fortran_g(123.456) == "    123.4560000     "
    assert fortran_g(-3.4187286) == "   -3.418728600     "
    assert fortran_g(0.0) == "    0.000000000     "
    assert fortran_g(1.0e-5) == "   0.1000000000E-004"
    assert fortran_g(-2.5e12) == "  -0.2500000000E+013"
    assert fortran_g(9999999999.4) == "    9999999999.     "
    assert fortran_g(0.09999999999) == "   0.9999999999E-001"
    assert fortran_g(float("nan")).strip() == "NaN"
    assert len(g_line([1.0, 2.0, 3.0])) == 61
    assert fortran_e(0.1, 20, 10, 3) == "   0.1000000000E+000"
    assert fortran_f(12.0, 7, 2) == "  12.00"
    assert fortran_g(5.25, 16, 8, 2) == "   5.2500000    "
    assert var(4, 2.0, 5.0) == 0.5


def test_checkpoint_and_rand_state_formats(tmp_path):
    rng = np.random.default_rng(0)
    P = rng.normal(size=(5, 3, 3))
    xe = rng.normal(size=(2, 3))
    f = tmp_path / "checkpoint.dat"
    write_checkpoint(f, False, P, xe, True, 2)
    trap, isopen, iworm, P2, xe2 = read_checkpoint(f, 3, 3, 2)
    assert (trap, isopen, iworm) == (False, True, 2)
    assert np.array_equal(P, P2) and np.array_equal(xe, xe2)          # 17 significant digits survive
    lines = open(f).read().splitlines()
    assert lines[0].strip() == ".False." and lines[1].strip() == ".True." and len(lines) == 3 + 15 + 2 + 2
    r = tmp_path / "rand_state"
    mt = rng.integers(0, 2 ** 32, 624, dtype=np.uint64).astype(np.uint32)
    append_rand_state(r, mt, 17)
    append_rand_state(r, mt[::-1].copy(), 99)          # the file only grows ...
    mt2, mti2 = read_rand_state(r)
    assert mti2 == 17 and np.array_equal(mt2, mt)      # ... and resume reads the OLDEST record (Q10)
    assert os.path.getsize(r) == 2 * (12 + 4 + 2496 + 4)


def test_driver_over_oracle_trap_zero_variance(tmp_path):
    """C1 (BASELINE configs[0]): non-interacting bosons in a trap -> the mixed estimator is exactly dim/2 per particle"""
    cfg = dict(C1, Nblock=2, Nstep=4, resume=False)
    d = VpiDriver(cfg, OracleBackend(cfg), workdir=str(tmp_path), potential="zero", quiet=True)
    res = d.run()
    assert res["diag_bl"] == 2 and abs(res["E"] - 1.5) < 1e-12
    rows = [ln.split() for ln in open(tmp_path / "e_vpi.out")]
    assert len(rows) == 2 and float(rows[0][0]) == 1.0 and abs(float(rows[1][1]) - 1.5) < 1e-9
    assert os.path.exists(tmp_path / "checkpoint.dat") and os.path.exists(tmp_path / "rand_state")
    assert not os.path.exists(tmp_path / "gr_vpi.out")            # structural estimators are PBC-only (vpi.f90:466)
    # resume: a second driver picks the state and the (oldest) random state up
    cfg2 = dict(cfg, resume=True, Nblock=1)
    d2 = VpiDriver(cfg2, OracleBackend(cfg2), workdir=str(tmp_path), potential="zero", quiet=True)
    assert abs(d2.run()["E"] - 1.5) < 1e-12


def test_driver_over_oracle_writes_all_files(tmp_path):
    cfg = dict(CWX, Nblock=3, Nstep=6)
    res = VpiDriver(cfg, OracleBackend(cfg), workdir=str(tmp_path), quiet=True).run()
    for f in ("e_vpi.out", "et_vpi.out", "gr_vpi.out", "sk_vpi.out", "jastrow.out", "potential.out", "fort.99",
              "checkpoint.dat", "rand_state"):
        assert os.path.getsize(tmp_path / f) > 0, f
    gr = np.loadtxt(tmp_path / "gr_vpi.out")
    assert gr.shape == (cfg["Nbin"], 3) and np.all(gr[:5, 1] == 0) and gr[-10:, 1].mean() > 0.3
    sk = np.loadtxt(tmp_path / "sk_vpi.out")
    assert sk.shape == (cfg["Nk"], 9)
    pot = np.loadtxt(tmp_path / "potential.out", max_rows=5000)
    assert pot.shape[1] == 2 and np.isnan(pot[0, 1])
    assert len(open(tmp_path / "fort.99").read().split()) == 2 * cfg["Np"]


@pytest.mark.gpu
@pytest.mark.parametrize("cfgname", ["CWX", "C1"])
def test_driver_gpu_replay_matches_driver_over_oracle(tmp_path, cfgname):
    from pathintegralgroundstate_b200 import PigsCuda
    cfg = dict(dict(CWX=CWX, C1=C1)[cfgname], Nblock=3, Nstep=8)
    pot = "zero" if cfgname == "C1" else "hfdb"
    a, b = tmp_path / "oracle", tmp_path / "gpu"
    VpiDriver(cfg, OracleBackend(cfg), workdir=str(a), potential=pot, quiet=True).run()
    sim = PigsCuda(cfg, n_chains=1, rng="mt", seed=cfg["seed"])
    VpiDriver(cfg, sim, workdir=str(b), potential=pot, quiet=True).run()
    files = ["e_vpi.out", "et_vpi.out", "fort.99", "checkpoint.dat"]
    if cfgname != "C1":
        files += ["gr_vpi.out", "sk_vpi.out"]
    for f in files:
        x = np.array([float(t) for t in open(a / f).read().replace(".True.", "1").replace(".False.", "0").split()])
        y = np.array([float(t) for t in open(b / f).read().replace(".True.", "1").replace(".False.", "0").split()])
        assert x.shape == y.shape, f
        ok = np.isclose(x, y, rtol=2e-7, atol=1e-9) | (np.isnan(x) & np.isnan(y))
        assert ok.all(), (f, x[~ok][:4], y[~ok][:4])
    # integer files are identical byte for byte; the random state record too
    assert open(a / "fort.99").read() == open(b / "fort.99").read()
    assert open(a / "rand_state", "rb").read() == open(b / "rand_state", "rb").read()


# ------------------------------------------------------------------ the compiled driver (vpi_cuda)
def test_vpi_cuda_g_editing_matches_python():
    rng = np.random.default_rng(7)
    xs = [0.0, -0.0, 1.0, 0.1, 0.09999999999, 0.099999999995, 9999999999.4, 9999999999.5, 1e10, 123.456, -3.4187286, 1e-5, -2.5e12,
          1e-310, 1.7976931348623157e308, 5e-324, 0.99999999995, 9.9999999995, 99999.999995, float("nan"), float("inf"), -float("inf")]
    xs += list(rng.normal(size=400) * 10.0 ** rng.integers(-12, 13, size=400))
    xs += [float(f"{m}e{e}") for m in ("9.99999999949", "9.9999999995", "9.99999999951", "1", "4.5", "-9.9999999995") for e in range(-4, 12)]
    cases = [(20, 10, 3, x) for x in xs] + [(16, 8, 2, x) for x in xs[:200]]
    out = _vpi(["--format-test"], "".join(f"{w} {d} {e} {float(x)!r}\n" for w, d, e, x in cases)).split("\n")[:-1]
    assert len(out) == len(cases)
    for (w, d, e, x), got in zip(cases, out):
        assert got == fortran_g(x, w, d, e), (x, got, fortran_g(x, w, d, e))


@pytest.mark.parametrize("cfgname", ["CWX", "C1"])
def test_vpi_cuda_reads_vpi_in_and_writes_the_tables(tmp_path, cfgname):
    cfg = dict(dict(CWX=CWX, C1=C1)[cfgname], Nblock=3, Nstep=8)
    pot = "zero" if cfgname == "C1" else "hfdb"
    text = "! a comment line\n" + format_vpi_in(cfg, cuda=dict(n_chains=7, rng="mt"))
    assert read_vpi_in(text)["cuda"] == dict(n_chains=7, rng="mt")
    a, b = tmp_path / "cxx", tmp_path / "py"
    out = _vpi(["--tables-only", "--workdir", str(a), "--potential", pot], text)
    kv = {ln.split()[0]: ln.split()[1:] for ln in out.splitlines()[1:]}
    g = derive_geometry(cfg)
    head = out.splitlines()[0].split()
    assert head[1::2][:4] == [str(cfg["dim"]), str(cfg["Np"]), str(cfg["Nb"]), "10000"] and head[9] == cfg["sampling"] and head[11] == "7"
    for k in ("rcut", "dr", "rbin", "density", "delta_cm"):
        assert float(kv[k][0]) == g[k], k                      # %.17g round-trips: identical doubles
    assert [float(t) for t in kv["Lbox"]] == list(g["Lbox"])
    VpiDriver(cfg, OracleBackend(cfg), workdir=str(b), potential=pot, quiet=True).tables()
    for f in ("jastrow.out", "potential.out"):
        assert open(a / f).read() == open(b / f).read(), f


def _compare_run_dirs(a, b, files):
    for f in files:
        assert os.path.exists(a / f) == os.path.exists(b / f), f
        if os.path.exists(a / f):
            assert open(a / f, "rb").read() == open(b / f, "rb").read(), f


@pytest.mark.gpu
@pytest.mark.parametrize("cfgname", ["CWX", "C1", "CS"])
def test_vpi_cuda_matches_python_driver_byte_for_byte(tmp_path, cfgname):
    """the compiled program and the Python driver are the same program over the same C ABI: every file identical"""
    from pathintegralgroundstate_b200 import PigsCuda
    from tests.common import CS
    cfg = dict(dict(CWX=CWX, C1=C1, CS=CS)[cfgname], Nblock=3, Nstep=8, checkpoint_every=0)     # reference-style files only
    pot = "zero" if cfgname == "C1" else "hfdb"
    a, b = tmp_path / "cxx", tmp_path / "py"
    out = _vpi(["--workdir", str(a), "--potential", pot], format_vpi_in(cfg, cuda=dict(n_chains=1, rng="mt", checkpoint_every=0)))
    d = VpiDriver(cfg, PigsCuda(cfg, n_chains=1, rng="mt", seed=cfg["seed"]), workdir=str(b), potential=pot, quiet=True)
    d.run()
    files = ["e_vpi.out", "et_vpi.out", "gr_vpi.out", "sk_vpi.out", "nr_vpi.out", "fort.99", "checkpoint.dat", "rand_state",
             "jastrow.out", "potential.out"]
    _compare_run_dirs(a, b, files)
    strip = lambda lines: [ln for ln in lines if not ln.startswith((" # Time per block", " # GPU throughput"))]
    assert strip(out.splitlines()) == strip(d.out)
    # resume from the files the compiled program wrote (vpi_mod.f90:162-185): both continue identically
    cfg2 = dict(cfg, resume=True, Nblock=2)
    _vpi(["--workdir", str(a), "--potential", pot], format_vpi_in(cfg2, cuda=dict(n_chains=1, rng="mt", checkpoint_every=0)))
    VpiDriver(cfg2, PigsCuda(cfg2, n_chains=1, rng="mt", seed=cfg["seed"]), workdir=str(b), potential=pot, quiet=True).run()
    _compare_run_dirs(a, b, ["e_vpi.out", "et_vpi.out", "checkpoint.dat", "fort.99"])


@pytest.mark.gpu
def test_vpi_cuda_many_chains_philox(tmp_path):
    cfg = dict(CWX, Nblock=2, Nstep=6)
    out = _vpi(["--workdir", str(tmp_path)], format_vpi_in(cfg, cuda=dict(n_chains=64, rng="philox")))
    assert "Markov chains (GPU) :    64" in out and "FINAL RESULTS" in out and " BLOCK NUMBER :           2" in out
    e = np.loadtxt(tmp_path / "e_vpi.out")
    assert e.shape == (2, 4) and np.all(np.isfinite(e))


@pytest.mark.gpu
def test_vpi_cuda_crystal_reads_config_ini(tmp_path):
    """crystal=T: Np, Lbox and density come from config_ini.in (vpi.f90:101-107), the sites from its tail (vpi_mod.f90:220-228)"""
    from pathintegralgroundstate_b200 import PigsCuda
    from pathintegralgroundstate_b200.host import write_config_ini
    from pathintegralgroundstate_b200.workloads import hcp_lattice
    R, L = hcp_lattice(2, 2, 2, density=0.48426)          # 32 hcp sites
    cfg = dict(CWX, Np=32, density=0.48426, crystal=True, Lbox=list(L), Nblock=2, Nstep=6)
    a, b = tmp_path / "cxx", tmp_path / "py"
    for d in (a, b):
        os.makedirs(d)
        write_config_ini(str(d / "config_ini.in"), R, L, 0.48426)
    vin = format_vpi_in(dict(cfg, Np=1, density=9.9), cuda=dict(n_chains=1, rng="mt"))     # Np/density in vpi.in are overridden
    out = _vpi(["--workdir", str(a)], vin)
    assert "Number of particles :    32" in out
    VpiDriver(cfg, PigsCuda(cfg, n_chains=1, rng="mt", seed=cfg["seed"]), workdir=str(b), quiet=True).run()
    _compare_run_dirs(a, b, ["e_vpi.out", "et_vpi.out", "gr_vpi.out", "sk_vpi.out", "fort.99", "checkpoint.dat", "rand_state"])
    # the Python command line reads config_ini.in the same way
    import subprocess, sys
    c = tmp_path / "pycli"
    os.makedirs(c)
    write_config_ini(str(c / "config_ini.in"), R, L, 0.48426)
    r = subprocess.run([sys.executable, "-m", "pathintegralgroundstate_b200.driver", "--workdir", str(c)], input=vin, capture_output=True,
                       text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    _compare_run_dirs(a, c, ["e_vpi.out", "et_vpi.out", "checkpoint.dat"])


@pytest.mark.gpu
@pytest.mark.parametrize("first,second", [("cxx", "cxx"), ("py", "py"), ("cxx", "py")])
def test_full_resume_continues_bit_for_bit(tmp_path, first, second):
    """the extended checkpoint (checkpoint_chains.bin: every chain's path, worm/permutation state and RNG;
    checkpoint_driver.bin: the accumulators the reference loses on restart): 2 blocks, stop, resume, 2 more blocks
    == 4 uninterrupted blocks, byte for byte in every output file -- also across the two drivers"""
    from pathintegralgroundstate_b200 import PigsCuda
    n = 6
    cuda = dict(n_chains=n, rng="philox")
    files = ["e_vpi.out", "et_vpi.out", "gr_vpi.out", "sk_vpi.out", "nr_vpi.out", "fort.99", "checkpoint.dat", "checkpoint_chains.bin",
             "checkpoint_driver.bin"]

    def run(which, wd, cfg):
        if which == "cxx":
            _vpi(["--workdir", str(wd)], format_vpi_in(cfg, cuda=cuda))
        else:
            VpiDriver(cfg, PigsCuda(cfg, n_chains=n, rng="philox", seed=cfg["seed"]), workdir=str(wd), quiet=True).run()

    whole, parts = tmp_path / "whole", tmp_path / "parts"
    run(first, whole, dict(CWX, Nblock=4, Nstep=8))
    run(first, parts, dict(CWX, Nblock=2, Nstep=8))
    assert np.loadtxt(parts / "e_vpi.out").shape[0] <= 2
    run(second, parts, dict(CWX, Nblock=4, Nstep=8, resume=True))
    if first != second:
        # the two drivers agree on every digit they print, but not on the last bit of every accumulator (numpy's power
        # vs std::pow in the ideal-gas normalisation): the binary accumulator file is compared as numbers there
        files = [f for f in files if f != "checkpoint_driver.bin"]
        a = np.frombuffer(open(whole / "checkpoint_driver.bin", "rb").read()[16:], "<f8")
        b = np.frombuffer(open(parts / "checkpoint_driver.bin", "rb").read()[16:], "<f8")
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=0)
    _compare_run_dirs(whole, parts, files)
    assert np.loadtxt(whole / "e_vpi.out").ndim == 2


NASTY_VPI_IN = """! leading comment with & and / characters
 &SYSTEM DIM=3, NP = 8 , Density = 1.0D0, TRAP = .TRUE. , crystal=.false. /
&samp
  resume = .FALSE. ! comment after value
  dt = 5.0d-2, Nb=20, seed = 7
  delta_cm = 0.5 , CMFreq = 1, sampling = "bis"
  Lstag = 16, Nlev = 3,
  Nstag = 5, Nblock = 2, Nstep = 3, Nbin = 10, Nk = 4
&end
&obdm swapping=T CWorm=0.25 Nobdm=1 Npw=0 /
&wavefun Nmax=1000, wf_table=T, v_table=T /
&jastrow
 Rm = 1.2 /
&extpot a_ho(1:3) = 1.0, 2.0d0 , 0.5 /
&cuda n_chains = 5, rng = 'MT', action = 'primitive' /
"""


def test_namelist_parsers_agree_on_free_form_input(tmp_path):
    """upper case, .TRUE./T, D exponents, '&end', inline comments, missing commas, index ranges, double quotes"""
    c = read_vpi_in(NASTY_VPI_IN)
    out = _vpi(["--tables-only", "--workdir", str(tmp_path), "--potential", "zero"], NASTY_VPI_IN).splitlines()
    head = out[0].split()
    assert head[1::2] == ["3", "8", "20", "1000", "bis", "5", "mt"]
    assert (c["dim"], c["Np"], c["Nb"], c["Nmax"], c["sampling"], c["cuda"]["n_chains"], c["cuda"]["rng"].lower()) == (3, 8, 20, 1000, "bis", 5, "mt")
    cfgline = [ln for ln in out if ln.startswith("config ")][0].split()[1:]
    got = dict(zip(cfgline[0::2], (int(v) for v in cfgline[1::2])))
    for k in ("resume", "seed", "CMFreq", "Lstag", "Nlev", "Nstag", "Nblock", "Nstep", "Nbin", "Nk", "swapping", "Nobdm", "Npw", "trap",
              "crystal", "wf_table", "v_table"):
        assert got[k] == int(c[k]), k
    reals = [ln for ln in out if ln.startswith("reals ")][0].split()
    assert float(reals[2]) == c["dt"] == 0.05 and float(reals[4]) == c["CWorm"] == 0.25 and float(reals[6]) == c["Rm"]
    assert [float(t) for t in reals[8:]] == c["a_ho"] == [1.0, 2.0, 0.5]
    assert out[-1] == "action primitive" and c["action"] == "primitive"
    g = derive_geometry(c)
    kv = {ln.split()[0]: ln.split()[1:] for ln in out[1:]}
    for k in ("rcut", "dr", "rbin", "density", "delta_cm"):
        assert float(kv[k][0]) == g[k], k
