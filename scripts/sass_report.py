#!/usr/bin/env python3
"""sass_report.py <object or .so> <kernel-substring> -> a SASS digest for profiles/: instruction-mix histogram of the
kernel, the evidence mnemonics (UBLKPF = TMA-engine bulk L2 prefetch, MUFU.RSQ64H, DFMA ...), and the full text of the
hot partner loop (the innermost backward branch with the most FP64 instructions)."""
import collections
import re
import subprocess
import sys

obj, name = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
# split per function
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur:
        funcs[cur].append(line)
key = [k for k in funcs if name in k][0]
ins = []
for l in funcs[key]:
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"kernel {key}: {len(ins)} SASS instructions ({len(ins) * 16 / 1024:.0f} KB)")
mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
print("instruction mix (whole kernel, static):")
for k, v in mix.most_common(28):
    print(f"  {k:12s} {v:6d}")
for ev in ("UBLKPF", "MUFU.RSQ64H", "DFMA", "LDG.E.NA.64", "BAR.SYNC", "LDS.64", "SHFL.BFLY"):
    print(f"evidence {ev:12s}: {sum(1 for _, t in ins if ev in t)}")
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA.*?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = ins[addr[tgt]:i + 1]
            inner = not any(re.search(r"BRA.*?0x([0-9a-f]+)", x[1]) and int(re.search(r"BRA.*?0x([0-9a-f]+)", x[1]).group(1), 16) < x[0]
                            and x[0] != a for x in body)
            nf = sum(1 for _, x in body if re.search(r"\bD(FMA|ADD|MUL|SETP)", x))
            ldg = sum(1 for _, x in body if "LDG" in x)
            if inner and ldg >= 3 and (best is None or nf > best[0]):
                best = (nf, body)
if best:
    nf, body = best
    print(f"\nhot partner loop: {len(body)} instructions, {nf} FP64, {sum(1 for _, x in body if 'LDS' in x)} LDS, "
          f"{sum(1 for _, x in body if 'LDG' in x)} LDG, {sum(1 for _, x in body if 'MUFU' in x)} MUFU (odd-slice and even/end-slice paths together)")
    for a, t in body:
        print(f"  /*{a:05x}*/ {t}")
