"""Per-kernel SASS opcode summary of the shipped libctcb200.so (cuobjdump -sass): which kernels issue Blackwell tensor-core
(UTCHMMA = tcgen05.mma), tensor-memory (LDTM / STTM = tcgen05.ld / .st), TMA (UTMALDG / UTMASTG / UTMAREDG) and bulk-copy
(UBLKCP = cp.async.bulk shared::cluster) instructions. Runs without a GPU.

usage: python tools/sass_summary.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ctc_pytorch_b200", "libctcb200.so")
out = sys.argv[1] if len(sys.argv) > 1 else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "MUFU", "FFMA", "HMMA",
        "LDS", "STS", "LDG", "STG", "ATOM", "RED", "SHFL", "BAR"]
kern = None
counts = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"ctcb200::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        kern = name
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[kern]["_total"] += 1
        for w in WANT:
            if op == w or op.startswith(w):
                counts[kern][w] += 1
                break
lines = ["# cuobjdump -sass ctc_pytorch_b200/libctcb200.so (sm_100a): static opcode counts per kernel",
         "# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/.st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce-add,",
         "# UBLKCP = cp.async.bulk (DSMEM copies with complete_tx), SYNCS = mbarrier ops", ""]
hdr = "%-64s %6s " % ("kernel", "instr") + " ".join("%8s" % w for w in WANT[:10])
lines.append(hdr)
for k, c in counts.items():
    lines.append("%-64s %6d " % (k[:64], c["_total"]) + " ".join("%8d" % c[w] for w in WANT[:10]))
text = "\n".join(lines) + "\n"
if out:
    open(out, "w").write(text)
print(text)
