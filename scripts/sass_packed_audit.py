#!/usr/bin/env python
"""Audit of the packed fp32 instructions in the production kernel's SASS (no GPU needed).

ptxas contracts `mul.rn.f32x2` followed by `add.rn.f32x2` into ONE FFMA2 even under -fmad=false (fastmath.cuh), which
would change torch's rounding.  The pass (lane_pass, beamopt_lanes.cuh) is written so that no packed product feeds a
packed sum; this script proves it for a build by COUNTING: per slot pair the source has exactly 28 fma2, 25 mul2 and
3 add2 calls (+ 3 packed adds per pair inside torch.sum's four-row block), so a contraction anywhere shows up as an
FFMA2 too many and an FMUL2 / FADD2 too few.  usage: python scripts/sass_packed_audit.py [lib.so]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "openpystruct_b200/lib/libopenpystruct_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ok = True
for m in re.finditer(r"Function : (\S*beamopt_lanes_kernelILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb([01])\S*)(.*?)(?=Function :|\Z)", sass, re.S):
    name, epl, nfix, nc, tfix, sc, body = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)), m.group(6), m.group(7)
    pairs = (epl + 1) // 2
    chain_pairs = pairs if nc == 1 else (pairs + nc - 1) // nc      # multi-case: only the OWNED pairs run the fp32 chain
    n = {op: len(re.findall(r"\b%s\b" % op, body)) for op in ("FFMA2", "FMUL2", "FADD2")}
    # a SCALAR fma that ptxas happens to issue in packed form (both multiplicands broadcast scalars, e.g. the Newton
    # step of 1 / sqrt(bias_correction2)) is not a pair operation of the pass
    n["FFMA2"] -= len(re.findall(r"FFMA2 R\d+, -?U?R\d+(?:\.reuse)?\.F32, -?U?R\d+(?:\.reuse)?\.F32,", body))
    want_fma = 28 * chain_pairs
    want_mul = 25 * chain_pairs if nc == 1 else 23 * chain_pairs  # (NC > 1: M^2, V^2 come from the exchange columns)
    # packed adds: 3 per pair in the Adam half; the torch.sum block adds depend on n (compile-time for nfix, else scalar)
    blk_pairs = (((nfix // 8) // 4) * 4) // 2 if nfix else 0
    want_add = 3 * chain_pairs + 3 * blk_pairs
    # a contraction turns one FMUL2 + one FADD2 into an FFMA2: the FMUL2 count is the proof (ptxas may ADD packed
    # instructions of its own -- two scalar adds of the generic-n sums as one FADD2, a scalar Newton step as an FFMA2 --
    # which changes no rounding)
    # An FMUL2 MORE than the source has (ptxas packing two scalar products, or cloning a block) cannot come from a
    # contraction; it is accepted only together with the exact FFMA2 count.
    extra_mul = n["FMUL2"] - want_mul
    good = extra_mul >= 0 and n["FADD2"] >= min(want_add, 3 * chain_pairs) and \
        (abs(n["FFMA2"] - want_fma) <= 1 if extra_mul == 0 else n["FFMA2"] == want_fma)
    ok &= good
    print(f"{'ok ' if good else 'BAD'} <EPL {epl:2d}, n {nfix:3d}, cases {nc}, T {tfix:3d}, scatter {sc}>  "
          f"FFMA2 {n['FFMA2']:3d} (want {want_fma})  FMUL2 {n['FMUL2']:3d} (want {want_mul})  FADD2 {n['FADD2']:3d} (want {want_add}{'+' if not nfix else ''})")
# the tensor-memory instances (beamopt_lanes_tm.cu: <EPL, n, T, scatter>, single case): the same pass
for m in re.finditer(r"Function : (\S*beamopt_lanes_tm_kernelILi(\d+)ELi(\d+)ELi(\d+)ELb([01])\S*)(.*?)(?=Function :|\Z)", sass, re.S):
    epl, nfix, tfix, sc, body = int(m.group(2)), int(m.group(3)), int(m.group(4)), m.group(5), m.group(6)
    pairs = (epl + 1) // 2
    n = {op: len(re.findall(r"\b%s\b" % op, body)) for op in ("FFMA2", "FMUL2", "FADD2")}
    n["FFMA2"] -= len(re.findall(r"FFMA2 R\d+, -?U?R\d+(?:\.reuse)?\.F32, -?U?R\d+(?:\.reuse)?\.F32,", body))
    want_fma, want_mul = 28 * pairs, 25 * pairs
    blk_pairs = (((nfix // 8) // 4) * 4) // 2 if nfix else 0
    want_add = 3 * pairs + 3 * blk_pairs
    extra_mul = n["FMUL2"] - want_mul
    good = extra_mul >= 0 and n["FADD2"] >= min(want_add, 3 * pairs) and \
        (abs(n["FFMA2"] - want_fma) <= 1 if extra_mul == 0 else n["FFMA2"] == want_fma)
    ok &= good
    print(f"{'ok ' if good else 'BAD'} tensor memory <EPL {epl:2d}, n {nfix:3d}, T {tfix:3d}, scatter {sc}>  "
          f"FFMA2 {n['FFMA2']:3d} (want {want_fma})  FMUL2 {n['FMUL2']:3d} (want {want_mul})  FADD2 {n['FADD2']:3d} (want {want_add}{'+' if not nfix else ''})")
print("packed audit:", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
