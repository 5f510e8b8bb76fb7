"""Reader for the reference's input file ``vpi.in`` (Fortran namelists on stdin).

Mirrors ``ReadParameters`` (vpi_mod.f90:14-80) and ``ReadSystemParameters``
(system_mod.f90:15-34): the groups ``&system &samp &obdm &wavefun &jastrow``
and, when ``trap=T``, ``&extpot``; defaults as in vpi_mod.f90:39-60.  A new
optional group ``&cuda`` carries what the reference does not have (n_chains,
rng, threads_per_chain, table_mode, gpus); the reference ignores unknown groups
because it looks every group up by name, so one file serves both programs.
"""
from __future__ import annotations

import re

DEFAULTS = dict(
    # &system
    crystal=False, trap=False,
    # &samp
    resume=False, seed=1982, Lstag=2, Nlev=1,
    # &obdm
    swapping=False, CWorm=0.0, Nobdm=0, Npw=0,
    # &wavefun
    Nmax=10000, wf_table=False, v_table=False,
)
GROUPS = dict(
    system=("dim", "Np", "density", "crystal", "trap"),
    samp=("resume", "dt", "Nb", "seed", "delta_cm", "CMFreq", "sampling", "Lstag", "Nlev", "Nstag", "Nblock",
          "Nstep", "Nbin", "Nk"),
    obdm=("swapping", "CWorm", "Nobdm", "Npw"),
    wavefun=("Nmax", "wf_table", "v_table"),
    jastrow=("Rm",),
    extpot=("a_ho",),
    cuda=("n_chains", "rng", "threads_per_chain", "table_mode", "gpus", "philox_seed", "action", "schedule", "checkpoint_every"),
)
REQUIRED = ("dim", "Np", "density", "dt", "Nb", "delta_cm", "CMFreq", "sampling", "Nstag", "Nblock", "Nstep", "Nbin",
            "Nk", "Rm")


def _value(tok: str):
    t = tok.strip()
    tl = t.lower().strip(".")
    if tl in ("t", "true"):
        return True
    if tl in ("f", "false"):
        return False
    if (t[:1] == t[-1:]) and t[:1] in "'\"" and len(t) >= 2:
        return t[1:-1]
    num = re.sub(r"[dD]", "e", t)
    try:
        if re.fullmatch(r"[+-]?\d+", num):
            return int(num)
        return float(num)
    except ValueError:
        return t


def parse_namelists(text: str) -> dict:
    """All ``&group ... /`` blocks of *text* as ``{group: {name: value}}`` (names keep the
    canonical spelling of GROUPS when known; Fortran names are case-insensitive)."""
    out = {}
    # strip comments (an '!' outside quotes ends the line)
    lines = []
    for ln in text.splitlines():
        q = None
        buf = []
        for ch in ln:
            if q:
                buf.append(ch)
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                buf.append(ch)
            elif ch == "!":
                break
            else:
                buf.append(ch)
        lines.append("".join(buf))
    body = "\n".join(lines)
    for m in re.finditer(r"&\s*(\w+)(.*?)(?:^\s*/|/\s*$|&end)", body, flags=re.S | re.M | re.I):
        g = m.group(1).lower()
        canon = {n.lower(): n for n in GROUPS.get(g, ())}
        vals = {}
        for am in re.finditer(r"(\w+)\s*(?:\(\s*[\d:,\s]*\))?\s*=\s*(.*?)(?=(?:,?\s*\w+\s*(?:\([^)]*\))?\s*=)|\Z)", m.group(2), flags=re.S):
            name = am.group(1)
            raw = am.group(2).strip().rstrip(",").strip()
            toks = [t for t in re.split(r"[,\s]+", raw) if t] if not (raw[:1] in "'\"") else [raw]
            v = [_value(t) for t in toks]
            vals[canon.get(name.lower(), name)] = v[0] if len(v) == 1 else v
        out[g] = vals
    return out


def read_vpi_in(text: str) -> dict:
    """The flat configuration the driver works with (variable names of vpi.in)."""
    nl = parse_namelists(text)
    cfg = dict(DEFAULTS)
    for g in ("system", "samp", "obdm", "wavefun", "jastrow"):
        cfg.update(nl.get(g, {}))
    if cfg.get("trap"):
        a = nl.get("extpot", {}).get("a_ho")
        if a is None:
            raise ValueError("trap=T needs &extpot a_ho (system_mod.f90:24-28)")
        cfg["a_ho"] = list(a) if isinstance(a, (list, tuple)) else [a]
    missing = [k for k in REQUIRED if k not in cfg]
    if missing:
        raise ValueError("vpi.in lacks: " + ", ".join(missing))
    if isinstance(cfg["sampling"], str):
        cfg["sampling"] = cfg["sampling"].strip()[:3]
    cfg["cuda"] = dict(nl.get("cuda", {}))
    if "action" in cfg["cuda"]:                    # 'chin' (the reference's live code) | 'primitive' (global_mod.f90:48,67)
        cfg["action"] = str(cfg["cuda"]["action"])
    return cfg


def format_vpi_in(cfg: dict, cuda: dict | None = None) -> str:
    """The inverse of read_vpi_in: a vpi.in for the reference (and, with the optional
    &cuda group, for vpi_cuda / this package's driver) from a flat configuration."""
    def lit(v):
        if isinstance(v, bool):
            return "T" if v else "F"
        if isinstance(v, int):
            return str(v)
        if isinstance(v, float):
            return repr(v).replace("e", "d") if "e" in repr(v) else repr(v) + "d0"
        if isinstance(v, (list, tuple)):
            return ", ".join(lit(x) for x in v)
        return "'" + str(v) + "'"
    full = dict(DEFAULTS)
    full.update(cfg)
    out = []
    for g in ("system", "samp", "obdm", "wavefun", "jastrow", "extpot"):
        if g == "extpot" and not full.get("trap"):
            continue
        out.append("&" + g)
        for name in GROUPS[g]:
            if name in full:
                out.append(f" {name} = {lit(full[name])},")
        out.append("/")
    cu = dict(cuda or cfg.get("cuda") or {})
    if cu:
        out.append("&cuda")
        for name in GROUPS["cuda"]:
            if name in cu:
                out.append(f" {name} = {lit(cu[name])},")
        out.append("/")
    return "\n".join(out) + "\n"
