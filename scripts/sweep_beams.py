#!/usr/bin/env python
"""Kernel time of the production path against the number of beams (GPU box): how an epoch's duration grows with the
resident warps per SM -- B = 148 * g beams puts g beams (g / 4 warps of the lanes kernel) on every SM.
usage: python scripts/sweep_beams.py [beams ...]   (env OPS_B200_LIB selects another build of the library)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                         # noqa: E402
from openpystruct_b200 import ops                                    # noqa: E402

counts = [int(a) for a in sys.argv[1:]] or [148 * g for g in (4, 8, 16, 24, 32, 40)] + [10000]
epochs = int(os.environ.get("SWEEP_EPOCHS", "600"))
early = os.environ.get("SWEEP_EARLY_STOP", "0") == "1"
bench.select_workload(os.environ.get("SWEEP_WORKLOAD", "cfg2"), 1)
bench.EPOCHS = epochs
p = bench.workload_params(early_stop=early)
dev = torch.device("cuda", 0)
fixed, fn, fv, L = bench.sample_inputs(max(counts), seed=int(os.environ.get("SWEEP_SEED", "1000")))
d_all = [torch.from_numpy(a).to(dev) for a in (fixed, fn, fv, L)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B in counts:
    d_in = [t[:B].contiguous() for t in d_all]
    for _ in range(3):
        out = ops.optimise_beams(p, *d_in)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ops.optimise_beams(p, *d_in); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    ep = float(out["epochs"].float().mean())
    print(f"B={B:8d} beams/SM={B / 148:7.2f} kernel_ms {med:8.3f} (min {ms[0]:.3f})  us_per_epoch {1e3 * med / ep:7.3f}  "
          f"beams/s {B / med * 1e3:10.0f}  mean_epochs {ep:.1f}", flush=True)
