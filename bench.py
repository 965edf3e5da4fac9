#!/usr/bin/env python
"""bench.py -- conv_fft output Gsamples/s on the BASELINE.json workload.

Workload (config.workload): BASELINE.json configs[4], the configuration the metric's "1/2/4/8 B200" and
"HBM GB/s % of peak" are quoted on, and the largest that fits one GPU:
    2-D conv_fft f32, x = (32768, 32768), k = (63, 63), ConvMode::Full, PaddingMode::Reflect  
    -> out (32830, 32830) = 1.078 G samples per step.
A step = one full convolution of the array.  With N > 1 ranks the OUTPUT rows of axis 0 are split into N
overlap-save slabs (ndconv_slab_plan); every rank convolves its slab (input rows + (Kd0-1)-row halo) with no
data-path collective -> strong scaling, total work fixed.

  value : whole-job output samples/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e   : same metric through the host-buffer C-ABI call (ndconv_conv_fft, NDCONV_MEM_HOST): pinned host input ->
          H2D -> kernels -> D2H -> pinned host output, all inside the timed region
  roofline : dominant kernel, algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the oracle's scipy-pocketfft restatement of the reference pipeline
          (good_size_cc buffers -> rfftn x2 -> multiply -> irfftn -> crop), all host cores, on a bounded row-slab
          of the same workload.  kind "port": the reference is Rust and cannot be built in this image.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path


def host_threads():
    """threads this process may run on (its CPU affinity mask, not the machine's core count)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


if "reference" in sys.argv[1:] or any(a.startswith("--impl=reference") for a in sys.argv[1:]):
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers when nproc > 1; pocketfft's pool and numpy's BLAS size
    # themselves from it at import time, which made the CPU arm ~4x slower at N >= 2 than at N = 1 (round-1 VERDICT).  The
    # reference arm always gets every host thread of its affinity mask, whatever launched it.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = str(host_threads())

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (x shape, k shape, dilation, mode, padding)
    "c5": dict(x=(32768, 32768), k=(63, 63), dil=1, mode="full", padding="reflect", dtype="float32",
               desc="2D conv_fft f32 x=(32768,32768) k=(63,63) Full Reflect (BASELINE configs[4])"),
    "c5s": dict(x=(8192, 9868), k=(63, 63), dil=1, mode="full", padding="reflect", dtype="float32",
                desc="reduced c5 for quick checks (NOT the headline)"),
}
# one tile-row stripes of c5 (experiments on L2 residency of the workspace; NOT bench lines)
for _n, _r in (("s1024", 900), ("s512", 388), ("s256", 132)):
    WORKLOADS[_n] = dict(x=(_r, 32768), k=(63, 63), dil=1, mode="full", padding="reflect", dtype="float32", desc=f"one {_n[1:]}-row tile stripe of c5 (experiment)")
WORKLOADS["r8"] = dict(x=(4104, 32768), k=(63, 63), dil=1, mode="full", padding="reflect", dtype="float32", desc="one rank's share of c5 on 8 GPUs (experiment)")
CPU_SAMPLE_ROWS = 4096   # rows of x in the bounded CPU sample


def metric_name():
    return "conv_fft output Gsamples/s"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi needs ~0.1 s to
    start, longer than a short timed region, so the sampler is started before the warm-up and the rows that arrive between
    mark_begin() and mark_end() are the ones reported."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.i0 = self.i1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark_begin(self):
        self.i0 = len(self.rows)

    def mark_end(self):
        self.i1 = len(self.rows)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)          # one more sampling period: the row that covers the end of the region
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        i0 = self.i0 if self.i0 is not None else 0
        i1 = (self.i1 if self.i1 is not None else len(self.rows)) + 1
        rows = self.rows[i0:i1] or self.rows[-1:]
        sm = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


GEN_BLOCK = 2048   # rows per generator block of the global synthetic array


def gen_rows(w, lo, hi, out=None):
    """Rows [lo, hi) of THE global synthetic input of workload `w`: Uniform[0,1) f32, generated per block of
    GEN_BLOCK rows from default_rng([1005, block]) so that any rank (and the reference arm) can materialise exactly its own
    rows of one and the same array -- the N-rank slabs assemble into the array the N = 1 run convolves."""
    n1 = int(np.prod(w["x"][1:]))
    if out is None:
        out = np.empty((hi - lo, n1), np.float32)
    for b in range(lo // GEN_BLOCK, (hi - 1) // GEN_BLOCK + 1):
        a, e = max(lo, b * GEN_BLOCK), min(hi, (b + 1) * GEN_BLOCK)
        blk = np.random.default_rng([1005, b]).random((GEN_BLOCK, n1), dtype=np.float32)
        out[a - lo:e - lo] = blk[a - b * GEN_BLOCK:e - b * GEN_BLOCK]
    return out


def gen_kernel(w):
    return np.random.default_rng(2005).random(w["k"], dtype=np.float32)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's scipy port of the reference pipeline on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_run(w, rows, steps, warmup, workers, budget_s):
    """`steps` timed repetitions (fewer if `budget_s` seconds run out; at least one) of the reference pipeline on the first
    `rows` input rows of the workload's global array.  Returns (Gsamples/s from the MEAN step, times, n_out, description)."""
    from oracle import oracle
    n0 = w["x"][0]
    rows = min(rows, n0)
    x, k = gen_rows(w, 0, rows), gen_kernel(w)
    times, out, t_start = [], None, time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = oracle.conv_fft_scipy(x, k, w["mode"], w["padding"], w["dil"], True, workers=workers)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        n_out = int(np.prod(out.shape))
        del out
        if times and time.perf_counter() - t_start + 1.5 * dt > budget_s:
            break
    full = rows == n0
    what = ("the full workload" if full else f"{rows} of {n0} input rows x {w['x'][1]} cols, same kernel/mode/border") + f" -> {n_out} output samples per step"
    return n_out / float(np.mean(times)) / 1e9, times, n_out, what, full


def cpu_sample(w, workers):
    """cpu_baseline leg of the GPU arm's line: ONE repetition on a bounded row slab (keeps the default run short)"""
    v, times, n_out, what, _ = cpu_run(w, CPU_SAMPLE_ROWS, 1, 0, workers, 60.0)
    return v, times[0], what


def host_mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1048576.0
    except Exception:
        pass
    return 0.0


def run_reference(args, w, rank, world):
    """The reference's own CPU implementation of the path (kind "port": scipy-pocketfft restatement of
    src/conv_fft/mod.rs:229-289; the Rust crate cannot be built in this image), every host thread, on the FULL workload --
    the same configuration the GPU arm's line is quoted on -- whenever the host has the memory for it (c5 holds ~40 GB live).
    A full c5 step takes several seconds, so the K requested steps are capped by a time budget; the line's `steps` / `warmup`
    are the repetitions actually run and `ms_per_step` is their MEAN (best is reported beside it).  Rank 0 only."""
    if rank != 0:
        return
    workers = host_threads()
    elems = int(np.prod(w["x"]))
    need_gb = elems * 4 * 9.5 / 2**30 + 2           # x, 2 padded buffers, 2 spectra, result (+ slack): c5 -> ~40 GB
    full_ok = host_mem_available_gb() >= need_gb and not args.ref_sample
    rows = w["x"][0] if full_ok else CPU_SAMPLE_ROWS
    warm = 1 if args.warmup else 0
    v, times, n_out, what, full = cpu_run(w, rows, max(1, args.steps), warm, workers, args.ref_budget_s)
    t_mean, t_best = float(np.mean(times)), float(np.min(times))
    line = {
        "impl": "reference", "metric": metric_name(), "value": v, "unit": "Gsamples/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": t_mean * 1e3, "ms_per_step_best": t_best * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": what, "same_config_as_gpu_arm": bool(full),
                   "sample_fraction": 1.0 if full else rows / w["x"][0],
                   "note": ("full workload; the requested steps are capped by a time budget of %.0f s" % args.ref_budget_s) if full else
                           ("host memory too small for the full workload (%.0f GB available, %.0f GB needed): bounded row slab, rate comparable, configuration not identical" % (host_mem_available_gb(), need_gb))},
        "cpu_baseline": {"value": v, "unit": "Gsamples/s", "cores": workers, "threads_used": workers, "kind": "port", "sample": what,
                         "engine": "scipy.fft (pocketfft, workers=%d) restatement of src/conv_fft/mod.rs:229-289; the Rust reference cannot be built here" % workers,
                         "omp_num_threads": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": v, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("ndarray-conv_b200")
    lib = pkg.get_library()
    # Which GPU a rank takes: with more GPUs visible than ranks, the ranks are spread evenly over the indices (N = 2 on an 8-GPU
    # box -> GPUs 0 and 4) instead of packing 0 .. N-1.  The host<->device legs of the e2e measurement share PCIe uplinks / host
    # memory ports in groups of neighbouring GPUs (profiles/r02_pcie_probe_8gpu.jsonl: two GPUs of one group 25.7 GB/s each way
    # per rank in duplex, one GPU from each group 34.3; four: 12.8 vs 17.8), so spreading is worth 1.3-1.4x end to end at N = 2 / 4.
    # Device-resident numbers do not depend on it.  --gpu-pick linear restores local_rank -> GPU local_rank.
    ngpu = torch.cuda.device_count()
    stride = ngpu // world if (args.gpu_pick == "spread" and world > 1 and ngpu >= 2 * world) else 1
    gpu_index = local_rank * stride
    torch.cuda.set_device(gpu_index)
    dev = torch.device("cuda", gpu_index)
    local_rank = gpu_index
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kshape, dil = w["k"], w["dil"]
    mode = {"full": pkg.ConvMode.Full, "same": pkg.ConvMode.Same, "valid": pkg.ConvMode.Valid}[w["mode"]]
    pmode = {"reflect": pkg.PaddingMode.Reflect, "zeros": pkg.PaddingMode.Zeros, "replicate": pkg.PaddingMode.Replicate,
             "circular": pkg.PaddingMode.Circular}[w["padding"]]
    n0, n1 = w["x"]
    k_host = gen_kernel(w)
    kwd = pkg.with_dilation(k_host, dil)

    # ---- slab of this rank (overlap-save along axis 0) ----
    pads, strides = mode.unfold(kshape, [dil] * 2, lib)
    # --as-slab W:R (kernel experiments on one GPU): time the slab rank R of a W-rank run would get; the line is NOT a bench value
    plan_world, plan_rank = (int(args.as_slab.split(":")[0]), int(args.as_slab.split(":")[1])) if args.as_slab else (world, rank)
    sl = pkg.slab_plan((n0, n1), np.float32, kwd, mode, pmode, pkg.PATH_FFT, plan_world, plan_rank, lib)
    Kd0 = (kshape[0] - 1) * dil + 1
    pf0, pb0 = int(pads[0][0]), int(pads[0][1])
    pb_, pe_ = sl["pad_begin"], sl["pad_end"]              # rows of the padded axis 0 this rank reads
    src_lo, src_hi = max(pb_ - pf0, 0), min(pe_ - pf0, n0)  # input rows held by this rank (halo included)
    slab_pf, slab_pb = max(pf0 - pb_, 0), max(pe_ - (pf0 + n0), 0)   # the true array edge falls inside the slab only at rank 0 / N-1
    rows = src_hi - src_lo
    explicit = ([[slab_pf, slab_pb], [int(pads[1][0]), int(pads[1][1])]], [int(strides[0]), int(strides[1])])
    out_rows = sl["out_end"] - sl["out_begin"]
    total_out = ((n0 + pf0 + pb0 - Kd0) // int(strides[0]) + 1) * ((n1 + int(pads[1][0]) + int(pads[1][1]) - ((kshape[1] - 1) * dil + 1)) // int(strides[1]) + 1)

    # ---- synthetic input: pinned host slab + device copy ----
    x_pin = torch.empty((rows, n1), dtype=torch.float32, pin_memory=True)
    xv = x_pin.numpy()
    gen_rows(w, src_lo, src_hi, xv)                       # this rank's rows (halo included) of the ONE global array
    x_dev = x_pin.to(dev, non_blocking=False)
    proc = pkg.get_fft_processor(local_rank, lib)
    # a dedicated (non-default) stream: kernels and the timing events live on the same stream; the legacy default
    # stream's handle is 0, which the C ABI reads as "use the processor's own stream"
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    proc.set_stream(stream.cuda_stream)
    oshape = pkg.conv_device("ndconv_conv_fft", proc, x_dev.data_ptr(), (rows, n1), (n1, 1), np.float32, kwd, mode, pmode, None, explicit=explicit)
    assert oshape[0] == out_rows, (oshape, out_rows)
    y_dev = torch.empty(oshape, dtype=torch.float32, device=dev)
    y_pin = torch.empty(oshape, dtype=torch.float32, pin_memory=True)

    prep_main = pkg.PreparedConv("ndconv_conv_fft", proc, (rows, n1), (n1, 1), np.float32, kwd, mode, pmode, explicit=explicit)
    xp_main, yp_main = x_dev.data_ptr(), y_dev.data_ptr()

    def step_device():
        prep_main(xp_main, yp_main)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- parity spot check against direct f64 evaluation (not timed) ----
    step_device()
    torch.cuda.synchronize(dev)
    spot = spot_check(xv, k_host, y_dev, slab_pf, int(pads[1][0]), w["padding"], dil, rows, n1)

    # ---- value: device-resident, CUDA events ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    clocks.mark_begin()
    l0 = proc.launch_count
    lib.c.ndconv_processor_set_profiling(proc.handle, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    clocks.mark_end()
    ms = e0.elapsed_time(e1)
    launches = proc.launch_count - l0
    kprof = read_profile(lib, proc)
    lib.c.ndconv_processor_set_profiling(proc.handle, 0)
    clk = clocks.stop() if rank == 0 else None
    t_dev = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_max = float(t_dev.item())

    shapes = other_shapes(pkg, lib, proc, dev, stream) if (rank == 0 and not args.no_shapes) else None

    # ---- e2e: host buffers through the C-ABI host call (H2D + kernels + D2H inside the timed region) ----
    proc.set_stream(0)   # the host call synchronises on the processor's own stream
    def step_host():
        pr, keep = pkg.make_problem((rows, n1), (n1, 1), x_pin.data_ptr(), np.float32, kwd, mode, pmode, pkg.MEM_HOST, lib, explicit=explicit)
        lib.check(lib.c.ndconv_conv_fft(proc.handle, pr, y_pin.data_ptr()))
    e2e_steps = max(1, min(args.steps, 5))
    if args.no_e2e:
        e2e_steps, t_e2e = 0, float("nan")
    else:
        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        barrier()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
    # ---- secondary: the same host call from ORDINARY (pageable) arrays, what a Vec-backed ndarray is (N = 1 only; not the headline) ----
    e2e_pageable = None
    if world == 1 and not args.no_e2e and not args.no_pageable:
        try:
            x_pg, y_pg = xv.copy(), np.zeros(tuple(oshape), np.float32)          # touched: no first-touch faults in the timed calls
            def step_pageable():
                pr, keep = pkg.make_problem((rows, n1), (n1, 1), x_pg.ctypes.data, np.float32, kwd, mode, pmode, pkg.MEM_HOST, lib, explicit=explicit)
                lib.check(lib.c.ndconv_conv_fft(proc.handle, pr, y_pg.ctypes.data))
            step_pageable()
            t0 = time.perf_counter()
            for _ in range(2):
                step_pageable()
            tp = (time.perf_counter() - t0) / 2
            same = bool(np.array_equal(y_pg[::257], y_pin.numpy()[::257]))
            e2e_pageable = {"value": total_out / tp / 1e9, "unit": "Gsamples/s", "ms_per_step": tp * 1e3, "steps": 2, "equals_pinned_result": same,
                            "note": "pageable numpy arrays through ndconv_conv_fft(NDCONV_MEM_HOST): slabs staged through pinned bounce buffers by host memcpy threads"}
            del x_pg, y_pg
        except Exception as exc:                                                    # a secondary figure must never take the bench line down
            e2e_pageable = {"error": repr(exc)[:200]}
    t_e = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e.item())
    # ---- the box's own floor for that leg: this step's bytes, H2D and D2H at the same time on two streams, no kernels, every rank
    # at once (what the host memory system / PCIe fabric of THIS box sustains; the e2e call cannot beat it) ----
    copy_floor = None
    if not args.no_e2e:
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        def copies():
            with torch.cuda.stream(s_in):
                x_dev.copy_(x_pin, non_blocking=True)
            with torch.cuda.stream(s_out):
                y_pin.copy_(y_dev, non_blocking=True)
        best = float("inf")
        for i in range(3):
            barrier()
            t0 = time.perf_counter()
            copies()
            torch.cuda.synchronize(dev)
            barrier()
            if i:
                best = min(best, time.perf_counter() - t0)
        t_c = torch.tensor([best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
        copy_floor = float(t_c.item())
    h2d = torch.tensor([x_pin.numel() * 4 + k_host.nbytes], dtype=torch.float64, device=dev)
    d2h = torch.tensor([y_pin.numel() * 4], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d)
        dist.all_reduce(d2h)

    # ---- N > 1: the N slabs, assembled, must be the one-GPU answer (same global array; not timed) ----
    assembled = verify_assembled(args, w, pkg, lib, proc, dev, stream, world, rank, kwd, mode, pmode, sl, y_dev,
                                 None if args.no_e2e else y_pin, total_out) if not args.no_verify else None

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = ms_max / args.steps
        value = total_out / (ms_step * 1e-3) / 1e9
        compulsory = 4.0 * (n0 * n1 + kshape[0] * kshape[1] + total_out)      # SURVEY 8d alg_bytes (whole job)
        roof = roofline_from_profile(kprof, args.steps, peak, peak_src, measured_traffic(args.workload) if world == 1 else None)
        cpu_v, cpu_t, cpu_s = cpu_sample(w, host_threads()) if world == 1 and not args.no_cpu else (None, None, None)
        line = {
            "metric": metric_name(), "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "out_samples_per_step": total_out, "parallelism": f"overlap-save slabs along axis 0 x{world}, no collective",
                       "gpu_pick": f"{args.gpu_pick}: rank r -> GPU {stride} r of {ngpu} visible",
                       "l2": "inputs (>= 0.5 GB per rank) are larger than the 126 MB L2; no explicit flush"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "e2e": {"value": total_out / t_e2e / 1e9, "unit": "Gsamples/s", "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": int(d2h.item()),
                    "ms_per_step": t_e2e * 1e3, "api": "ndconv_conv_fft(NDCONV_MEM_HOST) on pinned host buffers",
                    "copy_floor_ms": None if copy_floor is None else copy_floor * 1e3,
                    "frac_of_copy_floor": None if copy_floor is None else copy_floor / t_e2e,
                    "copy_floor_note": "the same bytes copied H2D and D2H at once by every rank with no kernels in between, measured in this run: the host<->device ceiling of this box for this step"},
            "e2e_pageable": e2e_pageable,
            "roofline": roof,
            "pipeline_compulsory": {"alg_bytes": compulsory, "achieved": compulsory / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": compulsory / (ms_step * 1e-3) / 1e9 / peak, "note": "SURVEY 8d compulsory bytes (in+kernel+out) over the whole step"},
            "kernels": kprof,
            "parity_spot_check": spot,
            "assembled_check": assembled,
            "workspace_bytes": proc.workspace_bytes,
        }
        if args.as_slab:
            line["invalid_as_slab"] = args.as_slab       # a kernel experiment: one slab of a larger run, not a bench value
        if shapes is not None:
            line["other_shapes"] = shapes
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": "Gsamples/s", "cores": host_threads(), "kind": "port", "sample": cpu_s, "ms": cpu_t * 1e3,
                                    "note": "one repetition on a bounded row slab; the full-workload figure is the --impl reference line"}
        print(json.dumps(line), flush=True)
    proc.close()
    if world > 1:
        dist.destroy_process_group()


def verify_assembled(args, w, pkg, lib, proc, dev, stream, world, rank, kwd, mode, pmode, sl, y_dev, y_pin, total_out):
    """Every rank's slab was cut from ONE global array (gen_rows).  Each rank contributes the float64 sum of each of its output
    rows and its first and last output row; rank 0 convolves the whole array on its own GPU (the N = 1 computation) and compares:
    all O0 row sums, and the 2N slab-edge rows element by element.  Also on every rank: the host-buffer (e2e) result against the
    device-resident one on every 257th row.  Tolerance: the float gate of the tests, 4 eps log2(F0 F1) max|out| (slabs pick their
    own tile lengths, so the two results differ by rounding)."""
    import torch
    import torch.distributed as dist
    n0, n1 = w["x"]
    O0 = total_out // y_dev.shape[1]
    O1 = y_dev.shape[1]
    eps = float(np.finfo(np.float32).eps)
    tol_rel = 4 * eps * np.log2(1024 * 2048)
    res = {}
    if y_pin is not None:
        a = y_pin[::257].to(dev, non_blocking=False)
        d = float((a - y_dev[::257]).abs().max())
        sc = float(y_dev[::257].abs().max())
        flag = torch.tensor([d / max(sc, 1e-30)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        res["e2e_vs_device_max_rel"] = float(flag.item())
        res["e2e_vs_device_ok"] = bool(float(flag.item()) <= 2 * tol_rel)
    if world == 1:
        return res or None
    rowsum = torch.zeros(O0, dtype=torch.float64, device=dev)
    rowsum[sl["out_begin"]:sl["out_end"]] = torch.sum(y_dev, dim=1, dtype=torch.float64)
    edges = torch.zeros((2 * world, O1), dtype=torch.float32, device=dev)
    edges[2 * rank] = y_dev[0]
    edges[2 * rank + 1] = y_dev[-1]
    dist.all_reduce(rowsum)
    dist.all_reduce(edges)
    if rank == 0:
        xf = torch.empty((n0, n1), dtype=torch.float32, device=dev)
        for lo in range(0, n0, 4 * GEN_BLOCK):
            hi = min(n0, lo + 4 * GEN_BLOCK)
            xf[lo:hi] = torch.from_numpy(gen_rows(w, lo, hi)).to(dev)
        yf = torch.empty((O0, O1), dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)                       # xf was filled on torch's stream; the processor runs on its own
        pkg.conv_device("ndconv_conv_fft", proc, xf.data_ptr(), (n0, n1), (n1, 1), np.float32, kwd, mode, pmode, yf.data_ptr())
        torch.cuda.synchronize(dev)
        ref = torch.sum(yf, dim=1, dtype=torch.float64)
        res["row_checksum_max_rel"] = float(((rowsum - ref).abs() / ref.abs().clamp_min(1e-30)).max())
        worst, scale = 0.0, float(yf.abs().max())
        for r in range(world):
            s_r = pkg.slab_plan((n0, n1), np.float32, kwd, mode, pmode, pkg.PATH_FFT, world, r, lib)
            worst = max(worst, float((edges[2 * r] - yf[s_r["out_begin"]]).abs().max()), float((edges[2 * r + 1] - yf[s_r["out_end"] - 1]).abs().max()))
        res["slab_edge_rows_max_abs_err"] = worst
        res["max_abs_out"] = scale
        res["tol_rel"] = tol_rel
        res["ok"] = bool(res["row_checksum_max_rel"] <= tol_rel and worst <= 2 * tol_rel * scale)
        res["what"] = f"{world} slabs of one global array vs the one-GPU convolution of the whole array on rank 0: {O0} float64 row sums + {2 * world} slab-edge rows"
        del xf, yf
    dist.barrier()
    return res


def other_shapes(pkg, lib, proc, dev, stream):
    """BASELINE configs[0..3] (parity-test shapes; latency-bound, microseconds): device-resident, warm processor,
    CUDA events around 50 back-to-back calls.  Secondary per-shape numbers, not the headline."""
    import torch
    B = pkg.BorderType
    cfgs = [
        ("c1 1D f32 x=5000 k=31 Same Zeros", "fft", np.float32, (5000,), (31,), 1, pkg.ConvMode.Same, pkg.PaddingMode.Zeros),
        ("c2 2D f32 x=(200,5000) k=(11,31) dil2 Same [Reflect,Circular]", "fft", np.float32, (200, 5000), (11, 31), 2, pkg.ConvMode.Same,
         pkg.PaddingMode.Custom([B.Reflect, B.Circular])),
        ("c3 3D f32 x=(10,100,200) k=(5,11,31) Same Zeros", "fft", np.float32, (10, 100, 200), (5, 11, 31), 1, pkg.ConvMode.Same, pkg.PaddingMode.Zeros),
        ("c3 3D Complex<f32> x=(10,100,200) k=(5,11,31) Same Zeros", "fft", np.complex64, (10, 100, 200), (5, 11, 31), 1, pkg.ConvMode.Same, pkg.PaddingMode.Zeros),
        ("c4 3D direct conv i32 x=(64,256,256) k=(3,5,5) pad(1,2,2) stride 2 Replicate", "direct", np.int32, (64, 256, 256), (3, 5, 5), 1,
         pkg.ConvMode.Custom([1, 2, 2], [2, 2, 2]), pkg.PaddingMode.Replicate),
        # not a BASELINE config: a large 1-D problem, where the rank-1 kernel is a single pass over memory (8 B per sample)
        ("extra 1D f32 x=67108864 k=63 Full Reflect", "fft", np.float32, (1 << 26,), (63,), 1, pkg.ConvMode.Full, pkg.PaddingMode.Reflect),
        # same-shape batches: a leading axis of kernel extent 1 is folded into one launch per pass (us_per_call / 64 = per problem)
        ("c2 x64 stacked: x=(64,200,5000) k=(1,11,31) dil (1,2,2) Same [Zeros,Reflect,Circular]", "fft", np.float32, (64, 200, 5000), (1, 11, 31), [1, 2, 2], pkg.ConvMode.Same,
         pkg.PaddingMode.Custom([B.Zeros, B.Reflect, B.Circular])),
        ("c3 x64 stacked: x=(64,10,100,200) k=(1,5,11,31) Same Zeros", "fft", np.float32, (64, 10, 100, 200), (1, 5, 11, 31), 1, pkg.ConvMode.Same, pkg.PaddingMode.Zeros),
        # the same shape in both precisions: f32 fast path (32 values per thread, packed FP32) and f64 fast path (16 values per thread; tiles of 512 x 256)
        ("extra 2D f32 x=(8192,8192) k=(63,63) Full Reflect (fast path)", "fft", np.float32, (8192, 8192), (63, 63), 1, pkg.ConvMode.Full, pkg.PaddingMode.Reflect),
        ("extra 2D f64 x=(8192,8192) k=(63,63) Full Reflect (f64 fast path)", "fft", np.float64, (8192, 8192), (63, 63), 1, pkg.ConvMode.Full, pkg.PaddingMode.Reflect),
        # the direct kernel in the throughput regime (not a BASELINE config): 75 multiply-adds per output make it ALU / shared-memory bound
        ("extra 3D direct conv i32 x=(256,1024,1024) k=(3,5,5) Same Replicate", "direct", np.int32, (256, 1024, 1024), (3, 5, 5), 1, pkg.ConvMode.Same, pkg.PaddingMode.Replicate),
    ]
    out = []
    for ci, (name, path, dt, xs, ks, dil, mode, pm) in enumerate(cfgs):
        rng = np.random.default_rng(1000 + ci)
        if dt == np.int32:
            xh = rng.integers(-128, 128, size=xs, dtype=np.int32)
            kh = np.random.default_rng(2000 + ci).integers(-128, 128, size=ks, dtype=np.int32)
        elif dt == np.complex64:
            xh = (rng.random(xs, dtype=np.float32) + 1j * rng.random(xs, dtype=np.float32)).astype(np.complex64)
            kr = np.random.default_rng(2000 + ci)
            kh = (kr.random(ks, dtype=np.float32) + 1j * kr.random(ks, dtype=np.float32)).astype(np.complex64)
        else:
            xh = rng.random(xs, dtype=np.float32).astype(dt)
            kh = np.random.default_rng(2000 + ci).random(ks, dtype=np.float32).astype(dt)
        xd = torch.from_numpy(xh.view(np.float32) if dt == np.complex64 else xh).to(dev)
        kw = pkg.with_dilation(kh, dil)
        entry = "ndconv_conv_fft" if path == "fft" else "ndconv_conv_direct"
        strides = [int(np.prod(xs[i + 1:])) for i in range(len(xs))]
        prep = pkg.PreparedConv(entry, proc, xs, strides, dt, kw, mode, pm)
        n_out = int(np.prod(prep.out_shape))
        yd = torch.empty(n_out * (2 if dt == np.complex64 else 1), dtype=torch.int32 if dt == np.int32 else (torch.float64 if dt == np.float64 else torch.float32), device=dev)
        xp, yp = xd.data_ptr(), yd.data_ptr()
        call = lambda: prep(xp, yp)
        for _ in range(5):
            call()
        torch.cuda.synchronize(dev)
        reps = 50
        l0 = proc.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            call()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) * 1e3 / reps
        nl = (proc.launch_count - l0) / reps
        read_profile(lib, proc)
        lib.c.ndconv_processor_set_profiling(proc.handle, 1)
        for _ in range(10):
            call()
        kp = read_profile(lib, proc)
        lib.c.ndconv_processor_set_profiling(proc.handle, 0)
        extra = {}
        if "stacked" in name:
            extra["us_per_problem"] = us / xs[0]
        if path == "direct":
            taps = int(np.count_nonzero(kh))
            extra["taps"] = taps
            extra["GMAC_per_s"] = n_out * taps / us / 1e3
        out.append({"shape": name, **extra, "us_per_call": us, "Gsamples_per_s": n_out / us / 1e3, "launches_per_call": nl,
                    "compulsory_GBps": (xh.nbytes + kh.nbytes + n_out * xh.itemsize) / us / 1e3,
                    "kernel_us_per_call": {k["kernel"]: round(k["total_ms"] * 1e3 / 10, 2) for k in kp},
                    "note": "device-resident, warm processor, includes host-side planning of every call"})
    return out


def spot_check(x_host, k, y_dev, pf0, pf1, padding, dil, rows, n1, n=48):
    """max |gpu - direct f64| over a few output samples of rank 0's slab, relative to max|out| (reflect / interior only)"""
    import torch
    rng = np.random.default_rng(1)
    O0, O1 = y_dev.shape
    Kd0, Kd1 = (k.shape[0] - 1) * dil + 1, (k.shape[1] - 1) * dil + 1
    kf = k[::-1, ::-1].astype(np.float64)
    worst, scale = 0.0, 0.0
    def src(i, pf, nn):
        j = i - pf
        if padding == "reflect":
            if j < 0:
                j = -j
            if j >= nn:
                j = 2 * (nn - 1) - j
        return j
    pts = [(0, 0), (O0 - 1, O1 - 1), (0, O1 - 1), (O0 - 1, 0)] + [(int(rng.integers(0, O0)), int(rng.integers(0, O1))) for _ in range(n)]
    yv = {p: float(y_dev[p[0], p[1]].item()) for p in pts}
    for (o0, o1), got in yv.items():
        r = [src(o0 + a * dil, pf0, rows) for a in range(k.shape[0])]
        c = [src(o1 + b * dil, pf1, n1) for b in range(k.shape[1])]
        if padding != "reflect" and (min(r) < 0 or max(r) >= rows or min(c) < 0 or max(c) >= n1):
            continue
        if min(r) < 0 or max(r) >= rows:
            continue   # slab-interior edge (halo rows are real neighbours; not reflected) -- skip
        ref = float(np.sum(x_host[np.ix_(r, c)].astype(np.float64) * kf))
        worst = max(worst, abs(got - ref))
        scale = max(scale, abs(ref))
    return {"points": len(pts), "max_abs_err": worst, "max_abs_out": scale, "rel": worst / max(scale, 1e-30)}


def read_profile(lib, proc):
    import ctypes
    names = (ctypes.c_char * 64 * 16)()
    ms = (ctypes.c_double * 16)()
    cnt = (ctypes.c_int64 * 16)()
    by = (ctypes.c_double * 16)()
    n = lib.c.ndconv_processor_get_profile(proc.handle, 16, names, ms, cnt, by)
    out = []
    for i in range(n):
        out.append({"kernel": bytes(names[i]).split(b"\0")[0].decode(), "launches": int(cnt[i]), "total_ms": float(ms[i]),
                    "alg_bytes_per_launch": float(by[i]) / max(int(cnt[i]), 1)})
    return out


def measured_traffic(workload):
    """per-launch DRAM traffic of each kernel from the latest committed ncu capture (profiles/r*_traffic_<workload>.json)"""
    cands = sorted((ROOT / "profiles").glob(f"r*_traffic_{workload}.json"))
    if not cands:
        return {}
    p = cands[-1]
    return {k: v["traffic_bytes"] for k, v in json.loads(p.read_text())["kernels"].items()}


def roofline_from_profile(kprof, steps, peak, peak_src, traffic=None):
    if not kprof:
        return None
    traffic = traffic or {}
    top = max(kprof, key=lambda k: k["total_ms"])
    avg_ms = top["total_ms"] / max(top["launches"], 1)
    ach = top["alg_bytes_per_launch"] / (avg_ms * 1e-3) / 1e9
    for k in kprof:
        a = k["total_ms"] / max(k["launches"], 1)
        k["avg_ms"] = a
        k["achieved_gbs"] = k["alg_bytes_per_launch"] / (a * 1e-3) / 1e9 if a > 0 else None
        k["frac_of_peak"] = k["achieved_gbs"] / peak if a > 0 else None
    for k in kprof:
        k["traffic_bytes_ncu"] = traffic.get(k["kernel"])
    return {"bound": "hbm", "kernel": top["kernel"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic.get(top["kernel"]),
            "peak_source": peak_src, "avg_launch_ms": avg_ms, "alg_bytes_per_launch": top["alg_bytes_per_launch"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pageable", action="store_true", help="skip the secondary pageable-array e2e leg")
    ap.add_argument("--no-e2e", action="store_true", help="kernel experiments only: skip the host-buffer leg (the line then has no valid e2e)")
    ap.add_argument("--no-shapes", action="store_true", help="skip the secondary per-shape numbers (configs c1-c4)")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the comparison of the assembled N-rank output with rank 0's one-GPU answer")
    ap.add_argument("--gpu-pick", default="spread", choices=["spread", "linear"], help="N ranks on a box with more GPUs: spread them over the GPU indices (default) or take GPUs 0 .. N-1")
    ap.add_argument("--as-slab", default="", help="W:R -- one GPU times the slab of rank R of W (per-kernel study of the N-GPU case; implies an invalid bench line)")
    ap.add_argument("--ref-sample", action="store_true", help="reference arm: force the bounded row-slab sample instead of the full workload")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="reference arm: wall-clock budget for its repetitions (at least one timed step is always run)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return
    run_gpu(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
