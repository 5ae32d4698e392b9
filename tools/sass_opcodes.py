"""SASS opcode histogram per kernel of the shipped library (profiles/r02_sass_opcodes.txt):
python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt   (cuobjdump -sass | c++filt, no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fourierflow_b200", "lib", "libffno_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "LDG", "STG", "LDS", "STS", "F2FP",
        "FFMA", "ELECT", "FENCE", "ACQBULK", "UTCATOMSWS", "ATOMG"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_n"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                per[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("SASS opcode histogram of the shipped library (cuobjdump -sass fourierflow_b200/lib/libffno_b200.so, sm_100a), per kernel.")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA tensor-map load),")
print("UBLKCP = cp.async.bulk (1-D bulk copy of weight / table images), UBLKPF = bulk L2 prefetch, LDGSTS = cp.async, SYNCS = mbarrier ops.")
print()
print("library total: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
print()
for (fn, c), name in sorted(zip(per.items(), names), key=lambda t: -t[0][1]["_n"]):
    print(name[:110])
    print(f"    {c['_n']} instructions: " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
