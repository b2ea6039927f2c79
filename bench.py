#!/usr/bin/env python
"""Benchmark of the hot path: optimised beams/s at a fixed epoch count (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA kernel behind the C ABI)
  python bench.py --impl reference [...]                       reference arm: the reference's CPU path
                                                               (torch-path port, all host cores)
  torchrun --nproc-per-node N bench.py --gpus N ...            one rank per GPU

A "step" is one pass of the fused optimisation loop over one batch of synthetic beams: BASELINE
configs[1] -- 10 000 beams per GPU, default discretisation (101 nodes, rollers 10/30/70/85/100, UDL
-1000, 1-4 point loads), max_e = 600 epochs with early stopping disabled (E_fix).
Inputs come from the seeded host sampler (the reference's draw order) and are resident in HBM before
the timed region; L2 is flushed between steps.  For N > 1 every rank optimises its own 10 000 beams
(weak scaling) and the per-step dataset all_gather over NCCL is inside the timed step.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "optimised beams/s (fixed 600 epochs)"
UNIT = "beams/s"
EPOCHS = 600
# BASELINE.json configs; the driver's default (and the only one `metric` is quoted on) is cfg2.  The others are
# selected with --workload for the extra lines kept under profiles/.
WORKLOADS = {
    "cfg2": dict(beams=10000, num_nodes=101, num_cases=1, rollers=None,
                 name="BASELINE configs[1]: BeamOpt_training_MultiCore dataset, 10k beams per GPU, default "
                      "discretisation, 600 fixed epochs (early stop off)"),
    "cfg3": dict(beams=1000000, num_nodes=101, num_cases=1, rollers=None, shard=True,
                 name="BASELINE configs[2]: GPU-batched generation, 1M beams sharded over the ranks, default "
                      "discretisation, 600 fixed epochs, dataset gather"),
    "cfg4": dict(beams=100000, num_nodes=101, num_cases=8, rollers=None, shard=True,
                 name="BASELINE configs[3]: MultiCase data, 8 load cases per beam sharing one I vector (summed "
                      "energies), 100k beams sharded over the ranks, 600 fixed epochs"),
    "cfg5": dict(beams=100000, num_nodes=1001, num_cases=1, rollers=[100, 300, 700, 850, 1000], shard=True,
                 name="BASELINE configs[4]: fine discretisation, 1000-element beams (rollers x10), 100k samples "
                      "sharded over the ranks, 600 fixed epochs"),
}
WL = WORKLOADS["cfg2"]
BEAMS_PER_GPU = WL["beams"]
NUM_NODES = WL["num_nodes"]
# algorithmic FP64 work per beam-iteration: assembly 8n + band LDL^T 16N + solves 13N +
# force recovery 16n = 82n + 58 with n elements, N = 2(n+1) DOFs (FMA = 2, div = sqrt = 1)
F64_FLOP_PER_ITER = 82 * (NUM_NODES - 1) + 58


def select_workload(name, world):
    """Rebinds the module-level workload constants (cfg2 unless --workload says otherwise)."""
    global WL, BEAMS_PER_GPU, NUM_NODES, F64_FLOP_PER_ITER, BYTES_PER_BEAM
    WL = WORKLOADS[name]
    NUM_NODES = WL["num_nodes"]
    BEAMS_PER_GPU = WL["beams"] // world if WL.get("shard") else WL["beams"]
    n, N, C = NUM_NODES - 1, 2 * NUM_NODES, WL["num_cases"]
    F64_FLOP_PER_ITER = 8 * n + 16 * N + C * (13 * N + 16 * n)          # SURVEY 8d (= 82 n + 58 for C = 1)
    BYTES_PER_BEAM = (NUM_NODES + C * 4 * 12 + 8) + (4 * n + C * (8 * n + 16 * NUM_NODES) + 12)
# algorithmic HBM bytes per beam: inputs (fixed_uy nn + force nodes/values 4*(4+8) + L 8) and
# outputs (I 4n, shear 4n, moment 4n, defl 8nn, rot 8nn, epochs/loss/status 12)
BYTES_PER_BEAM = (NUM_NODES + 4 * 12 + 8) + (12 * (NUM_NODES - 1) + 16 * NUM_NODES + 12)
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12      # 64 DFMA lanes/SM x 148 SMs x max SM clock
EXECUTED_F64_FLOP_PER_ITER = (226 * 2 + 67 + 66) * 32 // 4   # per beam-iteration, from the ncu counts (see roofline.executed_fp64)


SOLVER = 0


def workload_params(early_stop=False):
    from openpystruct_b200.params import BeamOptParams
    return BeamOptParams.for_script("MC").replace(early_stop=early_stop, max_e=EPOCHS, num_nodes=NUM_NODES,
                                                  num_cases=WL["num_cases"], solver=SOLVER)


def sample_inputs(beams, seed):
    """Seeded host sampling in the reference's draw order, by the native sampler (csrc/sampler_host.cpp:
    the stream of random.seed(seed), bit for bit)."""
    from openpystruct_b200 import generator
    p = workload_params()
    cfg = generator.GeneratorConfig(params=p, roller_nodes=tuple(WL["rollers"]) if WL["rollers"] else None)
    return generator.sample_cases_packed(cfg, beams * p.num_cases, seed).abi_arrays()


# --------------------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [x.strip() for x in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU baselines (the oracle is only ever the thing measured HERE, never on the product path)
# --------------------------------------------------------------------------------------------------
def cpu_torch_port(beams, workers, early_stop=False):
    """The reference's torch path (port) over a process pool, MultiCore:258 pattern: E_fix = 600 epochs, or the
    script's own early stopping (tolerance 5e-3, effective patience 10)."""
    from oracle import beamopt_port as port
    p = port.BeamOptParams.for_script("MC")
    p.early_stop = early_stop
    p.max_e = EPOCHS
    pool = port.PortPool(p, workers)
    try:
        pool.run(workers)                       # first call per worker pays the lazy torch / scipy initialisation
        done, dt = pool.run(beams, seed=1234)
    finally:
        pool.close()
    return done / dt, done, dt


def cpu_c_oracle(beams, workers):
    """The plain-C restatement on `workers` threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from tests.helpers import oracle_run
    p = workload_params()
    fixed, fn, fv, L = sample_inputs(beams, seed=4321)
    chunks = np.array_split(np.arange(beams), workers)
    oracle_run(p, fixed[:1], fn[:1], fv[:1], L[:1])        # build + warm
    t0 = time.perf_counter()
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(lambda c: oracle_run(p, fixed[c], fn[c], fv[c], L[c]) if len(c) else None, chunks))
    dt = time.perf_counter() - t0
    return beams / dt, dt


def measure_config(name, beams, steps, tf_peak):
    """Kernel-only throughput of another BASELINE config on this GPU (inputs resident, CUDA events, L2 flushed between
    steps), for the `configs` block of the default line."""
    import torch
    from openpystruct_b200 import ops
    saved = (WL, BEAMS_PER_GPU, NUM_NODES, F64_FLOP_PER_ITER, BYTES_PER_BEAM)
    keep = WORKLOADS[name]["beams"]
    try:
        WORKLOADS[name]["beams"] = beams
        select_workload(name, 1)
        p = workload_params()
        dev = torch.device("cuda", torch.cuda.current_device())
        d_in = [torch.from_numpy(a).to(dev) for a in sample_inputs(beams, seed=2000)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(2):                      # (the host-side sampling before this left the GPU idle: two warm-up launches)
            out = ops.optimise_beams(p, *d_in)
        torch.cuda.synchronize()
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = ops.optimise_beams(p, *d_in); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        times = [a.elapsed_time(b) for a, b in evs]
        ms = statistics.median(times)
        assert int(out["epochs"].min()) == EPOCHS and int(out["status"].sum()) == 0
        tf = beams * EPOCHS * F64_FLOP_PER_ITER / (ms * 1e-3) / 1e12
        res = {"workload": WL["name"], "beams": beams, "num_nodes": NUM_NODES, "num_cases": WL["num_cases"],
               "value": beams / (ms * 1e-3), "unit": UNIT, "kernel_ms": ms, "kernel_ms_steps": [round(t, 3) for t in times],
               "flop_per_beam_iteration": F64_FLOP_PER_ITER, "roofline_frac": tf / tf_peak}
        del out, d_in, flush
        torch.cuda.empty_cache()
        return res
    finally:
        WORKLOADS[name]["beams"] = keep
        globals().update(dict(zip(("WL", "BEAMS_PER_GPU", "NUM_NODES", "F64_FLOP_PER_ITER", "BYTES_PER_BEAM"), saved)))


def measure_frames(steps):
    """Frame optimiser: frames/s for a batch of random frames at a fixed epoch count."""
    import random
    import torch
    from openpystruct_b200 import frames
    p = frames.FrameOptParams(num_epochs=300, early_stop=False)
    rng = random.Random(0)
    batch = [frames.draw_frame(p, rng) for _ in range(592)]
    dev = torch.device("cuda", torch.cuda.current_device())
    nb = torch.tensor([f[0] for f in batch], dtype=torch.int32, device=dev)
    ns = torch.tensor([f[1] for f in batch], dtype=torch.int32, device=dev)
    frames.optimise_frames_device(p, nb, ns)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = frames.optimise_frames_device(p, nb, ns)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert int(out["status"].sum()) == 0 and int(out["epochs"].min()) == 300
    return {"workload": "OpenPyStruct_FrameOpt_Discrete_Beta.py: 592 random frames (1-10 bays x 1-10 stories), 300 fixed "
                        "epochs, one CTA per frame", "frames": len(batch), "value": len(batch) / (ms * 1e-3), "unit": "frames/s",
            "kernel_ms": ms, "frame_epochs_per_s": len(batch) * 300 / (ms * 1e-3)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import beamopt_port as port
    cores = host_cores()
    beams_per_step = max(cores * 2, 8)
    p = port.BeamOptParams.for_script("MC")
    p.early_stop = False
    p.max_e = EPOCHS
    pool = port.PortPool(p, cores)
    try:
        for _ in range(args.warmup):
            pool.run(max(cores, 4))
        times = []
        for _ in range(max(args.steps, 1)):
            done, dt = pool.run(beams_per_step)
            assert done == beams_per_step
            times.append(dt)
    finally:
        pool.close()
    value = beams_per_step * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: BeamOpt_training_MultiCore dataset, default discretisation, "
                               "600 fixed epochs; bounded sample per step", "beams_per_step": beams_per_step,
                   "num_nodes": NUM_NODES, "epochs": EPOCHS, "early_stop": False},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{beams_per_step} beams x {EPOCHS} epochs per step on a {cores}-process pool: "
                                   "reference torch path (torch.sum/autograd/Adam/ExponentialLR on CPU) with the "
                                   "OpenSees half restated (scipy dpbsv); OpenSeesPy is not installable offline"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _launch_plan(p, B):
    """Kernel family and launch geometry of the timed launch on this device (ops_beamopt_plan)."""
    from openpystruct_b200 import _cabi
    try:
        return _cabi.launch_plan(p, B, 0, 0)
    except Exception as e:                                  # a diagnostic must not cost the bench line
        return {"error": str(e)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from openpystruct_b200 import _cabi, ops
    from openpystruct_b200.distributed import PeerDataset, PeerUnavailable, gather_outputs, init_from_env

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank, local, world = init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if args.beams > 0:
        WORKLOADS[args.workload]["beams"] = args.beams
        WORKLOADS[args.workload]["name"] += f" [beam count overridden: {args.beams}]"
    select_workload(args.workload, world)
    global SOLVER
    SOLVER = args.solver
    p = workload_params()
    B = BEAMS_PER_GPU
    t_s0 = time.perf_counter()
    fixed, fn, fv, L = sample_inputs(B, seed=1000 + rank)
    sampling_s = time.perf_counter() - t_s0
    h_in = [torch.from_numpy(a).pin_memory() for a in (fixed, fn, fv, L)]
    d_in = [t.to(dev) for t in h_in]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # N > 1: every rank optimises its own beams; the dataset gather is fused into the kernel's record write
    # (rows copied to every peer's dataset arrays over NVLink, PeerDataset) unless --gather nccl asks for the
    # all_gather of the per-rank blocks after the kernel
    peer = None
    if world > 1 and args.gather == "peer" and NUM_NODES <= 169 and args.solver == 0:     # the lanes kernel's scatter
        try:
            peer = PeerDataset(p, B * world)
            shard = dict(zip(("fixed_uy", "force_nodes", "force_vals", "L"), d_in))
        except PeerUnavailable as ex:                 # agreed on by all ranks: NCCL gather instead
            if rank == 0:
                print(f"bench.py: {ex}; using the NCCL gather", file=sys.stderr, flush=True)

    def step():
        if peer is not None:
            return peer.optimise(shard, rank * B, check=False)       # (launch status checked once after the timed steps)
        out = ops.optimise_beams(p, *d_in)
        if world > 1:
            out = gather_outputs(out, B * world)
        return out

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    gather_check = None
    if peer is not None:
        # Driver-visible identity of the in-kernel dataset gather (one warm-up step, outside the timed region): the
        # dataset the kernels scattered into THIS rank's arrays over NVLink must equal, byte for byte, the NCCL
        # all_gather of the per-rank results -- on every rank, or the run fails loudly.
        got = {k_: v_.clone() for k_, v_ in step().items()}
        want = gather_outputs(ops.optimise_beams(p, *d_in), B * world)
        bad = [k_ for k_ in want if not torch.equal(got[k_].view(torch.uint8), want[k_].view(torch.uint8))]
        flag_t = torch.tensor([len(bad)], dtype=torch.int32, device=dev)
        dist.all_reduce(flag_t, op=dist.ReduceOp.MAX)
        if int(flag_t.item()) != 0:
            raise SystemExit(f"bench.py: rank {rank}: the in-kernel dataset gather differs from the NCCL gather in {bad}")
        gather_check = {"in_kernel_gather_equals_nccl_all_gather": True, "ranks": world, "rows": int(B * world),
                        "bytes_compared_per_rank": int(sum(v_.numel() * v_.element_size() for v_ in want.values()))}
        del got, want

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    if peer is not None:
        peer.check()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    assert int(out["epochs"].min()) == EPOCHS and int(out["status"].sum()) == 0

    # kernel-only duration (no gather), for the roofline of the dominant kernel
    kev = []
    ops.optimise_beams(p, *d_in)                 # (first call of this path in the peer mode: allocations)
    torch.cuda.synchronize()
    for _ in range(min(args.steps, 5)):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.optimise_beams(p, *d_in); e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)

    # reference-default mode for orientation: MultiCore's early stopping (tolerance 5e-3, patience 10)
    p_es = workload_params(early_stop=True)
    ops.optimise_beams(p_es, *d_in)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out_es = ops.optimise_beams(p_es, *d_in)
    e1.record()
    torch.cuda.synchronize()
    es_ms = e0.elapsed_time(e1) / 3
    es_epochs = float(out_es["epochs"].float().mean())

    # end to end through the C ABI with HOST buffers: every step copies that step's inputs from pinned host
    # memory to the device, runs the loop and copies the whole record (I, u, theta, V, M, epochs, loss,
    # status) back to pinned host memory (ops_beamopt_session_run; buffers allocated once, like a
    # generator that produces batch after batch)
    e2e_steps = max(args.steps, 3)
    sess = _cabi.Session(p, B, device=local)
    sess.load(fixed, fn, fv, L)
    for _ in range(2):
        sess.run(B)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sess.inputs["L"][:B] = L                      # the caller refreshes an input in place every step
        host_out = sess.run(B)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(e2e_s.item())
    assert int(host_out["epochs"].min()) == EPOCHS and int(host_out["status"].sum()) == 0
    assert np.array_equal(host_out["I"], out["I"][:B].cpu().numpy()) if world == 1 else True
    h2d = sum(a.nbytes for a in (fixed, fn, fv, L))
    d2h = sum(v.nbytes for v in host_out.values())
    sess.close()
    # one-shot variant (allocates, copies from pageable memory, frees): ops_beamopt_run_host
    _cabi.run_host(p, fixed, fn, fv, L, device=local)
    t0 = time.perf_counter()
    _cabi.run_host(p, fixed, fn, fv, L, device=local)
    oneshot_value = B / (time.perf_counter() - t0)

    # the call a user makes: generate_columnar = native sampling (reference draw order) + pinned session + columnar record
    # arrays, every step from a fresh seed; sampling, H2D, kernel, D2H and the record assembly all inside the timed region
    api = None
    if args.workload == "cfg2" and world == 1:
        from openpystruct_b200 import generator
        gcfg = generator.GeneratorConfig(params=p)
        for _ in generator.stream_columnar(gcfg, num_samples=2 * B, batch_size=B, seed=1, reuse_buffers=True):
            pass
        t0 = time.perf_counter()
        recs = 0
        for col in generator.stream_columnar(gcfg, num_samples=e2e_steps * B, batch_size=B, seed=100, reuse_buffers=True):
            recs += int(len(col["L"]))
            checksum = float(col["I_values"][-1, -1])                  # (touch the batch before it is recycled)
        api = {"value": recs / (time.perf_counter() - t0), "unit": UNIT,
               "path": "generator.stream_columnar(seed): batches of one seeded stream -- native sampler (reference draw "
                       "order, next batch drawn on a host thread while the GPU runs) -> ops_beamopt_session_run -> columnar "
                       "record arrays on the host (views of the pinned buffers, valid until the next batch)",
               "records": recs, "batch": B}
        t0 = time.perf_counter()
        col = generator.generate_columnar(gcfg, num_samples=B, seed=7)
        api["one_call_generate_columnar"] = B / (time.perf_counter() - t0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B * args.steps / (total_ms * 1e-3)
    # FP64 roofline of the dominant kernel (algorithmic flops only; DDIV expansion etc. not credited)
    tf_measured, _ = _cabi.fp64_peak_probe(1 << 16, torch.cuda.current_stream().cuda_stream)
    achieved_tf = B * EPOCHS * F64_FLOP_PER_ITER / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_achieved = B * BYTES_PER_BEAM / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        if args.workload == "cfg2":     # the capture is of the default workload's kernel
            traffic = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass

    configs = None
    if world == 1 and args.workload == "cfg2" and not args.no_configs:
        # the other BASELINE configs and the frame optimiser on this GPU (kernel-only, few steps): driver-visible
        configs = {"cfg3": measure_config("cfg3", 1000000, 3, tf_measured),
                   "cfg4": measure_config("cfg4", 100000, 3, tf_measured),
                   "cfg5": measure_config("cfg5", 100000, 3, tf_measured),
                   "frames": measure_frames(2)}

    cores = host_cores()
    cpu = None
    cpu_c = None
    if not args.no_cpu_baseline and world == 1 and args.workload == "cfg2":
        sample_beams = max(cores * 13, 208)                      # >= 200 beams
        rate, done, dt = cpu_torch_port(sample_beams, cores)
        es_rate, es_done, es_dt = cpu_torch_port(sample_beams, cores, early_stop=True)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{done} beams x {EPOCHS} fixed epochs in {dt:.1f} s on a {cores}-process pool "
                         "(reference torch path port: real torch.sum/autograd/Adam on CPU + scipy dpbsv for "
                         "the OpenSees half; OpenSeesPy not installable offline)",
               "early_stop_mode": {"value": es_rate, "unit": UNIT,
                                   "sample": f"{es_done} beams with the MultiCore script's early stopping (tolerance 5e-3, "
                                             f"patience 10) in {es_dt:.1f} s, same pool"}}
        c_beams = max(cores * 150, 600)
        c_rate, c_dt = cpu_c_oracle(c_beams, cores)
        cpu_c = {"value": c_rate, "unit": UNIT, "cores": cores, "kind": "port",
                 "sample": f"{c_beams} beams x {EPOCHS} fixed epochs in {c_dt:.1f} s, plain-C restatement "
                           f"(oracle/csrc) on {cores} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if WL.get("shard") else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WL["name"],
                   "beams_per_gpu": B, "num_nodes": NUM_NODES, "num_cases": WL["num_cases"], "epochs": EPOCHS,
                   "early_stop": False, "launch_plan": _launch_plan(p, B),
                   "fe_precision": "f64", "optimiser_precision": "f32 (torch CPU op order)",
                   "l2": "flushed between steps (256 MiB write)", "collective": "none" if world == 1 else (
                       "dataset gather fused into the kernel: every beam's record is copied to all peers' dataset "
                       "arrays over NVLink by the thread group that finished it (CUDA IPC mappings); NCCL carries "
                       "two 4-byte all_reduce barriers per step" if peer is not None else
                       "NCCL all_gather of the dataset after the kernel, every step")},
        "clocks": clocks,
        "gather_check": gather_check,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "path": "ops_beamopt_session_run (C ABI; pinned host buffers, H2D of the inputs + launch + D2H of the "
                        "whole record inside every step)",
                "one_shot_run_host": oneshot_value, "api_e2e": api},
        "configs": configs,
        "gpu_launches": args.steps,
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": tf_measured, "unit": "TFLOP/s",
                     "frac": achieved_tf / tf_measured, "traffic": traffic,
                     "peak_source": "measured in this run by ops_fp64_peak_probe (pure DFMA kernel); "
                                    f"nominal {NOMINAL_FP64_TFLOPS:.1f} TFLOP/s = 148 SM x 64 DFMA/clk x 1.965 GHz; "
                                    "MEASURED_PEAKS.json has no FP64 entry",
                     "frac_of_nominal": achieved_tf / NOMINAL_FP64_TFLOPS,
                     # what the three-moment kernel actually EXECUTES (ncu instruction counts of the production instance,
                     # profiles/r02_*: 226 DFMA + 67 DMUL + 66 DADD warp instructions per 4-beam warp-epoch, padding lanes
                     # included) -- the credited figure above is SURVEY 8d's band-LDL^T count, per the contract
                     "executed_fp64": {"flop_per_beam_iteration": EXECUTED_F64_FLOP_PER_ITER,
                                       "tflops": B * EPOCHS * EXECUTED_F64_FLOP_PER_ITER / (kernel_ms * 1e-3) / 1e12,
                                       "frac": B * EPOCHS * EXECUTED_F64_FLOP_PER_ITER / (kernel_ms * 1e-3) / 1e12 / tf_measured}
                     if args.workload in ("cfg2", "cfg3") else None,
                     "kernel": {"cfg2": "beamopt_lanes_kernel<13,100,1>", "cfg3": "beamopt_lanes_kernel<13,100,1>",
                                "cfg4": "beamopt_lanes_kernel<13,0,8>", "cfg5": "beamopt_wide_kernel<32>"}[args.workload],
                     "kernel_ms": kernel_ms,
                     "flop_per_beam_iteration": F64_FLOP_PER_ITER,
                     "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_achieved / hbm_peak, "bytes_per_beam": BYTES_PER_BEAM,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}},
        "early_stop_mode": {"value": B / (es_ms * 1e-3), "unit": UNIT, "kernel_ms": es_ms, "mean_epochs": es_epochs,
                            "note": "same beams with the MultiCore script's early stopping (tolerance 5e-3, patience "
                                    "10) instead of 600 fixed epochs; this rank only"},
        "host_side": {"sampling_and_packing_s": sampling_s,
                      "note": "Python `random` sampling of this rank's beams in the reference's draw order + packing "
                              "into the ABI arrays; outside every timed region (SURVEY 7: host-side costs are reported "
                              "separately)"},
        "cpu_baseline": cpu,
        "cpu_baseline_c": cpu_c,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 / frames sub-results")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg2")
    ap.add_argument("--beams", type=int, default=0, help="override the workload's total beam count (exploration only)")
    ap.add_argument("--gather", choices=["peer", "nccl"], default="peer",
                    help="N > 1: dataset gather fused into the kernel (stores to the peers over NVLink) or NCCL all_gather")
    ap.add_argument("--solver", type=int, default=0, help="OPS_SOLVER_* of the C ABI (exploration only; 0 = production)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
