#!/usr/bin/env python3
"""Aggregate an ncu source-page CSV (SASS level) by CUDA source line using nvdisasm line info.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all obj.o ; nvdisasm -g -c x.cubin > lines.txt
    python tools/ncu_by_line.py src.csv lines.txt <mangled-kernel-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict


def main(src_csv, lines_txt, kern, top=45):
    # 1. offsets -> (file, line) of the OUTERMOST non-libm frame we can see (nvdisasm prints the innermost
    #    location first and 'inlined at' frames after it)
    loc = {}
    cur = None
    infn = False
    for ln in open(lines_txt):
        if ln.startswith("//---") and ".text." in ln:
            infn = kern in ln
            cur = None
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            f, l, rest = m.group(1), int(m.group(2)), m.group(3)
            if "inlined at" in rest and cur is not None and not cur[0].endswith((".cuh", ".cu")):
                # previous was a library frame: replace by this caller frame
                m2 = re.search(r'inlined at "([^"]+)", line (\d+)', rest)
            if f.endswith((".cuh", ".cu")):
                cur = (f.split("/")[-1], l)
            else:
                m2 = re.search(r'inlined at "([^"]+\.cuh?)", line (\d+)', rest)
                cur = (m2.group(1).split("/")[-1], int(m2.group(2))) if m2 else (f.split("/")[-1], l)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m and cur is not None:
            loc[int(m.group(1), 16)] = (cur, m.group(2))
    rows = list(csv.reader(open(src_csv)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    iA, iS, iN, iI = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    base = min(int(r[iA], 16) for r in data)
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    opc = defaultdict(int)
    for r in data:
        off = int(r[iA], 16) - base
        inst, smp = int(r[iI]), int(r[iN])
        k = loc.get(off, (("?", 0), ""))[0]
        a = agg[k]
        a[0] += inst; a[1] += smp; a[2] += 1
        tot_i += inst; tot_s += smp
        op = r[iS].split()[0] if not r[iS].strip().startswith("@") else r[iS].split()[1]
        opc[op.split(".")[0]] += inst
    print(f"total warp-instructions executed {tot_i}, samples {tot_s}, static instrs {len(data)}")
    print("--- by source line (sorted by stall samples) ---")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[0]:>18}:{k[1]:<5} inst {a[0]:>10} ({100*a[0]/tot_i:5.1f}%)  samples {a[1]:>7} ({100*a[1]/max(tot_s,1):5.1f}%)  static {a[2]}")
    print("--- dynamic opcode mix ---")
    for op, n in sorted(opc.items(), key=lambda kv: -kv[1])[:30]:
        print(f"{op:>14} {n:>11} {100*n/tot_i:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 45)
