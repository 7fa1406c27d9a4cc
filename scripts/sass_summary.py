"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (profiles/r02_sass_summary.txt):
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = legacy mma.sync, FFMA2 / FADD2 = packed fp32x2.   usage: python scripts/sass_summary.py"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "t2v_turbo_b200", "libt2v_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "HMMA", "FFMA2", "FADD2", "MUFU", "REDG", "RED"]
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("t2v::", "")
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        for k in MN:
            if op == k or (k == "RED" and op == "RED"):
                cur[k] += 1
        cur["_total"] += 1
agg = collections.OrderedDict()
for name, c in per.items():
    fam = re.sub(r"<.*", "", name)
    a = agg.setdefault(fam, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}: SASS mnemonic counts per kernel family (all template instantiations summed)")
print(f"{'kernel':34s} {'inst':>4s} " + " ".join(f"{k:>7s}" for k in MN) + f" {'total':>8s}")
tot = collections.Counter()
for fam, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][1]["_total"]):
    print(f"{fam[:34]:34s} {n:4d} " + " ".join(f"{c[k]:7d}" for k in MN) + f" {c['_total']:8d}")
    tot.update(c)
print(f"{'ALL':34s} {sum(n for n, _ in agg.values()):4d} " + " ".join(f"{tot[k]:7d}" for k in MN) + f" {tot['_total']:8d}")
