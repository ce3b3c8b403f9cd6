"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md: UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, UBLKRED = cp.reduce.async.bulk, LDGSTS = cp.async, HMMA = legacy mma.sync), from the in-tree
libcomet_b200.so.  No GPU needed:  python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "comet_pose_estimation_b200", "libcomet_b200.so")
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "UTCBAR", "SYNCS", "HMMA", "HGMMA", "LDGSTS"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", name)
        counts.setdefault(cur, collections.Counter())
        counts[cur]["_instr"] += 0
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["_instr"] += 1
        for k in MN:
            if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
                counts[cur][k] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.basename(lib)} (cuobjdump -sass, sm_100a)")
print(f"# {'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in MN))
tot = collections.Counter()
for name, c in counts.items():
    tot.update(c)
    if any(c[k] for k in MN if k != "SYNCS") or c["_instr"] > 2000:
        print(f"{name[:72]:72s} {c['_instr']:7d} " + " ".join(f"{c[k]:8d}" for k in MN))
print(f"{'TOTAL (' + str(len(counts)) + ' kernels)':72s} {tot['_instr']:7d} " + " ".join(f"{tot[k]:8d}" for k in MN))
