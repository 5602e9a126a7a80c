"""Per-kernel counts of the SASS mnemonics that prove the sm_100a paths (tcgen05 MMA = UTCHMMA, TMEM loads = LDTM, tensor-memory
copies = UTCCP, TMA loads / stores / reductions = UTMALDG / UTMASTG / UTMAREDG, bulk copies = UBLKCP, mbarrier = SYNCS) in the
built library.  Run here (no GPU needed): python scripts/sass_summary.py > profiles/sass_summary_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rcnn-ocr_b200", "librcnn_ocr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
mn = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCCP", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "LDGSTS", "MUFU",
      "HMMA", "LDL", "STL", "MEMBAR", "CCTL"]
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("rcnn::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        if op in mn:
            counts[cur][op] += 1
print("sm_100a SASS of rcnn-ocr_b200/librcnn_ocr_b200.so (cuobjdump -sass): instruction counts per kernel")
print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{m:>8s}" for m in mn))
for k in order:
    c = counts[k]
    print(f"{k[:64]:64s} {c['_total']:6d} " + " ".join(f"{c[m]:8d}" if c[m] else f"{'.':>8s}" for m in mn))
