"""Aggregate an ncu source page (csv) by source line: samples, instructions, top stall."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# find header rows ("Line No" first col); multiple file sections
agg = collections.defaultdict(lambda: collections.Counter())
cur_file = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    # the SASS view rows carry 'Address'; source rows carry line text. Use rows with a source line number
    try: ln = int(d["Line No"])
    except: continue
    key = (cur_file, ln, r[1][:90])
    def f(k):
        try: return float(d.get(k) or 0)
        except: return 0.0
    a = agg[key]
    a["samples"] += f("# Samples"); a["inst"] += f("Instructions Executed")
    for k in ("stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math", "stall_lg", "stall_barrier", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_mio", "stall_dispatch"):
        a[k] += f(k)
    a["local"] += f("L2 Theoretical Sectors Local")
tot = sum(a["samples"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print("total samples", tot, "total inst", toti)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:n]:
    st = sorted(((v, k) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:3]
    print(f"{100*a['samples']/tot:5.1f}% smp {100*a['inst']/toti:5.1f}% inst  {key[0]}:{key[1]:4d} {' '.join(f'{k[6:]}={100*v/max(a['samples'],1):.0f}%' for v,k in st)} | {key[2].strip()[:70]}")
