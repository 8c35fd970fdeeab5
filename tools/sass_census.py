"""SASS census of libBridge.so: per kernel, how many TMA (UTMALDG / UTMASTG / UBLKCP), mbarrier (SYNCS), cp.async (LDGSTS),
128-bit global / shared accesses and FP64 instructions it contains.  python tools/sass_census.py > profiles/sass_census_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rvdd-release_b200", "lib", "libBridge.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "ELECT", "LDGSTS", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "SHFL", "MUFU",
        "DFMA", "DADD", "DMUL", "F2F", "BAR.SYNC", "ATOMG", "RED"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in keys:
            if op == k or op.startswith(k + ".") or (("." in k) and op.startswith(k)):
                counts[cur][k] += 1
print("# cuobjdump -sass %s (sm_100a)" % os.path.relpath(lib, ROOT))
print("%-58s %7s  %s" % ("kernel", "instrs", "selected mnemonics"))
for k, c in counts.items():
    print("%-58s %7d  %s" % (k[:58], total[k], ", ".join("%s %d" % kv for kv in c.items() if kv[1])))
