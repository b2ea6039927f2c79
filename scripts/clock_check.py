#!/usr/bin/env python
"""SM clock and power while the production kernel runs back to back for ~2 s at several batch sizes (GPU box)."""
import os, subprocess, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from openpystruct_b200 import ops
bench.select_workload("cfg2", 1)
p = bench.workload_params()
dev = torch.device("cuda", 0)
fixed, fn, fv, L = bench.sample_inputs(23680, seed=1000)
d_all = [torch.from_numpy(a).to(dev) for a in (fixed, fn, fv, L)]
for B in [int(a) for a in sys.argv[1:]] or [592, 2368, 5920, 10000, 23680]:
    d_in = [t[:B].contiguous() for t in d_all]
    ops.optimise_beams(p, *d_in); torch.cuda.synchronize()
    rows = []
    proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "50"],
                            stdout=subprocess.PIPE, text=True)
    th = threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True); th.start()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(20):
            ops.optimise_beams(p, *d_in); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    proc.terminate()
    ms = e0.elapsed_time(e1) / n
    clk = [float(r.split(",")[0]) for r in rows if r and r.split(",")[0].strip().replace('.','').isdigit()]
    pw = [float(r.split(",")[1]) for r in rows if r and len(r.split(",")) > 1 and r.split(",")[1].strip().replace('.','').isdigit()]
    clk = clk[len(clk)//3:]; pw = pw[len(pw)//3:]
    print(f"B={B:6d} back-to-back ms/launch {ms:7.3f}  sm clock MHz min/med/max {min(clk):.0f}/{sorted(clk)[len(clk)//2]:.0f}/{max(clk):.0f}  power W med {sorted(pw)[len(pw)//2]:.0f} max {max(pw):.0f}  reasons {set(r.split(',')[2].strip() for r in rows[len(rows)//3:] if len(r.split(','))>2)}", flush=True)
