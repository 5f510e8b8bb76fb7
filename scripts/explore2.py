"""Tuning matrix on the GPU box: each case in its own process (PIGS_LIB / PIGS_PREFETCH are read at import/create)."""
import json, os, subprocess, sys
cases = json.loads(sys.argv[1]) if len(sys.argv) > 1 else []
for c in cases:
    env = dict(os.environ)
    if c.get("lib"): env["PIGS_LIB"] = os.path.abspath(c["lib"])
    env["PIGS_PREFETCH"] = str(c.get("pf", 1))
    if "pfd" in c: env["PIGS_PFDIST"] = str(c["pfd"])
    r = subprocess.run([sys.executable, "scripts/prof_case.py", c["cfg"], str(c["n"]), str(c["T"]), str(c["tm"]), str(c.get("nstep", 2)), "1"],
                       env=env, capture_output=True, text=True)
    print(json.dumps(c), "=>", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr[-300:], flush=True)
