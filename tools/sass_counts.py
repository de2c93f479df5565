"""SASS mnemonic counts per kernel of blah2_b200/lib/libb200dd.so (cuobjdump -sass): the evidence that the TMA bulk
copies (UBLKCP), the mbarrier waits (SYNCS), the packed FP32 forms (FFMA2 / FADD2 / FMUL2) and the FP64 FMAs are what
the kernels execute.   python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "blah2_b200", "lib", "libb200dd.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "DFMA", "DADD", "DMUL", "F2F", "LDS", "STS", "LDG", "STG", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, mix = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        mix[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        mix[cur][m.group(1)] += 1
dem = subprocess.run(["c++filt"] + list(mix), capture_output=True, text=True).stdout.splitlines()
print(f"{'kernel':70s} total  " + " ".join(f"{w:>7s}" for w in WATCH))
for (name, c), d in sorted(zip(mix.items(), dem), key=lambda t: t[1]):
    short = re.sub(r"\(anonymous namespace\)::", "", d)
    short = re.sub(r"\(.*\)$", "", short).replace("void ", "")
    if len(sys.argv) > 1 and sys.argv[1] not in short:
        continue
    print(f"{short[:70]:70s} {sum(c.values()):5d}  " + " ".join(f"{c.get(w, 0):7d}" for w in WATCH))
tot = collections.Counter()
for c in mix.values():
    tot.update(c)
print(f"{'ALL KERNELS':70s} {sum(tot.values()):5d}  " + " ".join(f"{tot.get(w, 0):7d}" for w in WATCH))
