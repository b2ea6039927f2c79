#!/usr/bin/env python
"""Static issue-cost estimate of the lanes kernel from its SASS (no GPU needed).

For every instruction of one kernel instance: the stall field (bits 105..108 of the encoding =
cycles before the same warp may issue again) attributed to the source line nvdisasm reports
(-lineinfo build), summed per function of beamopt_lanes.cuh.  Sum(stall) over a phase is the
single-warp issue time of that phase when no scoreboard wait intervenes -- the quantity ptxas'
scheduling (interleaving of independent chains) decides.  Usage:
    python scripts/sass_stalls.py [lib.so] [kernel-substring]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openpystruct_b200/lib/libopenpystruct_b200.so")
kern = sys.argv[2] if len(sys.argv) > 2 else "beamopt_lanes_kernelILi13ELi100ELi1"
src = os.path.join(ROOT, "openpystruct_b200/csrc/beamopt_lanes.cuh")

# function line ranges of the header
funcs = []
for i, line in enumerate(open(src), 1):
    m = re.match(r"OPS_HD\s+[\w:<>\s\*&]+?\s+(\w+)\s*\(", line)
    if m:
        funcs.append((i, m.group(1)))
# line range of the cold generic-operator branch of the pass (Adam's v below the fast square root's range)
_src = open(src).read().split("\n")
_g0 = next(i for i, t in enumerate(_src, 1) if "const float dx = sqrtf(" in t)
GENERIC = (_g0 - 1, _g0 + 4)


def func_of(line):
    name = "?"
    for start, f in funcs:
        if start <= line + 1:
            name = f
    return name

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if "lanes" in f and f.endswith(".cubin") and "kernels-" not in f][0]
    text = subprocess.run(["nvdisasm", "-gi", "-hex", os.path.join(td, cubin)], capture_output=True, text=True).stdout

sec = text.split("//--------------------- .text.")
body = [s for s in sec if kern in s.split("\n", 1)[0]][0].split("\n")
cur_line, cur_file = 0, ""
in_line, in_file, chain_open = 0, "", False
agg = collections.OrderedDict()
i = 0
while i < len(body):
    ln = body[i]
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not chain_open:                      # first line of a chain = innermost frame
            in_file, in_line = os.path.basename(m.group(1)), int(m.group(2))
            chain_open = True
        if "inlined at" not in ln:              # last line of a chain = outermost frame (the kernel)
            cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
            chain_open = False
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", ln)
    if m and i + 1 < len(body):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", body[i + 1])
        if m2:
            stall = (int(m2.group(1), 16) >> 41) & 0xF
            inner = func_of(in_line) if in_file == "beamopt_lanes.cuh" else in_file.replace(".cuh", "")
            if inner == "lane_pass":
                inner = "lane_pass(generic ops)" if GENERIC[0] <= in_line <= GENERIC[1] else "lane_pass"
            key = (cur_file, cur_line, inner)
            a = agg.setdefault(key, [0, 0, 0])
            a[0] += 1; a[1] += stall; a[2] += stall >= 4
            i += 2
            continue
    i += 1
cu = open(os.path.join(ROOT, "openpystruct_b200/csrc/beamopt_lanes.cu")).read().split("\n")
print(f"{'call site (outermost line)':60s} {'instrs':>7s} {'sum stall':>9s} {'avg':>5s} {'stall>=4':>8s}")
tot = [0, 0]
for (f_, l_, inner), (n, s, f) in sorted(agg.items(), key=lambda kv: (kv[0][1], kv[0][2])):
    if n < 12:
        continue
    txt = cu[l_ - 1].strip()[:30] if f_ == "beamopt_lanes.cu" and 0 < l_ <= len(cu) else f_
    txt = f"{txt} > {inner}"
    print(f"{l_:4d} {txt[:55]:55s} {n:7d} {s:9d} {s / n:5.2f} {f:8d}")
    tot[0] += n; tot[1] += s
print(f"{'total (listed)':60s} {tot[0]:7d} {tot[1]:9d} {tot[1] / tot[0]:5.2f}")
