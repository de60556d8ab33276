#!/usr/bin/env python
"""bench.py -- measures BASELINE.json's headline metric on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: the batched C2C f32
FFT of BASELINE configs[1] (N = 4096 x batch 65536, 2 GiB in + 2 GiB out) through
`kofft_cuda_fft_c2c_f32` (one kernel launch).  Each rank owns one GPU and processes its own
65536 rows (rows are independent: no data-path collective; weak scaling).

Printed JSON line (rank 0):
  value      whole-job GFLOP/s (5 N log2 N per transform), inputs resident in HBM, CUDA-event timed
  e2e        the same metric through the host-pointer C ABI call (`kofft_cuda_fft_batch_host_f32`,
             what `CudaFftImpl::batch` binds) with pinned HOST buffers: H2D + kernel + D2H per step
  roofline   algorithmic bytes (16 N B per launch) / measured launch time vs the
             measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port of kofft's CPU path timed on this box's host cores (bounded sample)
  extra      STFT (configs[3] shape) frames/s and rfft numbers measured after the headline region

`--impl reference` times the CPU restatement of the reference (oracle/, "port": the Rust crate
cannot be built in this image) on the same config with all host threads.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import math
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 4096
ROWS_PER_GPU = 65536
FLOPS_PER_TRANSFORM = 5 * N * int(math.log2(N))
ALGO_BYTES_PER_TRANSFORM = 16 * N  # read 8N + write 8N
METRIC = "batched_c2c_f32_fft_gflops"
UNIT = "GFLOP/s"


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": f"batched C2C f32 FFT N={N} x batch {ROWS_PER_GPU} per GPU (BASELINE configs[1]), "
                    "uniform[-1,1) synthetic rows, out-of-place",
        "n": N,
        "rows_per_gpu": ROWS_PER_GPU,
        "global_batch": ROWS_PER_GPU * n_gpus,
        "parallelism": f"batch-sharded x{n_gpus}, no collective",
        "l2_policy": "inputs larger than L2 (2 GiB in + 2 GiB out per step vs 126 MB L2)",
        "arithmetic": "exact (bit-identical to the reference's unfused f32 arithmetic)",
    }


def peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic() -> float | None:
    """dram bytes per launch of the headline kernel from the committed ncu summary, if any."""
    path = os.path.join(ROOT, "profiles", "headline_kernel_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if bit and (mask & bit):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self) -> dict:
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
def cpu_port_gflops(rows: int, reps: int, threads: int) -> tuple[float, float]:
    """kofft's CPU path (oracle port) on `rows` transforms of length N, all `threads` cores:
    rows split evenly over threads, one reused planner per thread.  Returns (GFLOP/s, seconds)."""
    from oracle import kofft_oracle as ko

    ko.build()
    rng = np.random.default_rng(0)
    x = (rng.uniform(-1, 1, (rows, N)) + 1j * rng.uniform(-1, 1, (rows, N))).astype(np.complex64)
    ko.fft_batch_inplace(x[: max(threads, 1)].copy(), nthreads=threads)  # warm the threads / tables
    t0 = time.perf_counter()
    for _ in range(reps):
        ko.fft_batch_inplace(x, nthreads=threads)
    dt = time.perf_counter() - t0
    return FLOPS_PER_TRANSFORM * rows * reps / dt / 1e9, dt


def cpu_port_stft_frames_per_s(threads: int, fresh_planner: bool) -> tuple[float, str]:
    """kofft's CPU STFT (oracle port) on a bounded sample of configs[3]: 64 channels x 60 s of 48 kHz audio,
    Hann 2048 / hop 512, channels split over `threads` cores.  fresh_planner = a new planner + twiddle table
    per frame, which is what `stft::parallel` does (src/stft.rs:260); False = one shared table (fair)."""
    from oracle import kofft_oracle as ko

    ko.build()
    ch, length, hop, win = 64, 2_880_000, 512, 2048
    nframes = -(-length // hop)
    rng = np.random.default_rng(2)
    sig = rng.uniform(-1, 1, (ch, length)).astype(np.float32)
    w = ko.hann(win)
    ko.stft_batch(sig[:, :48_000], w, hop, -(-48_000 // hop), fresh_planner=fresh_planner, nthreads=threads)  # warm-up
    t0 = time.perf_counter()
    ko.stft_batch(sig, w, hop, nframes, fresh_planner=fresh_planner, nthreads=threads)
    dt = time.perf_counter() - t0
    return ch * nframes / dt, f"{ch} ch x {length} samples ({ch * nframes} frames, {dt:.1f} s), {threads} threads over channels"


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = 4096
    for _ in range(max(args.warmup, 1)):
        cpu_port_gflops(rows, 1, threads)
    t0 = time.perf_counter()
    vals = [cpu_port_gflops(rows, 1, threads)[0] for _ in range(args.steps)]
    total = time.perf_counter() - t0
    value = statistics.median(vals)
    sample = f"{rows} of {ROWS_PER_GPU} rows per step (N={N}), {threads} threads, rows split evenly; scaled by flops"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * FLOPS_PER_TRANSFORM * rows / (value * 1e9),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "kofft is a Rust crate and cannot be built in this image (no rustc/cargo); this arm times the "
                "bit-faithful C restatement of its CPU path (oracle/kofft_oracle.c) on the host cores",
        "wall_s": total,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    import kofft_b200
    from kofft_b200 import stft as S
    from kofft_b200 import window as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kofft_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at VERSION/INFO level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fft = kofft_b200.CudaFftImpl(device=local, exact=True)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.view_as_complex((torch.rand((ROWS_PER_GPU, N, 2), generator=g, device=dev) * 2 - 1)).contiguous()
    y = torch.empty_like(x)

    # ---- device-resident timing: K launches bracketed by events on the launching stream -------
    for _ in range(max(args.warmup, 3)):
        fft.fft_batch(x, out=y)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = fft.ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        fft.fft_batch(x, out=y)
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop()
    launches = fft.ctx.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = FLOPS_PER_TRANSFORM * ROWS_PER_GPU * n_gpus / (ms_per_step * 1e-3) / 1e9

    avg_launch_ms = sum(per_launch_ms) / len(per_launch_ms)
    peak, peak_src = peaks()
    achieved = ALGO_BYTES_PER_TRANSFORM * ROWS_PER_GPU / (avg_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "peak_source": peak_src,
                "kernel": "fft_cta_kernel<12, exact, IoC2C<false>>",
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_TRANSFORM * ROWS_PER_GPU,
                "launch_ms_avg": avg_launch_ms, "launch_ms_min": min(per_launch_ms),
                "frac_of_8TBs_spec": achieved / 8000.0}

    # ---- end to end through the host-pointer C ABI (pinned host buffers) -----------------------
    e2e_steps = max(2, min(args.steps, 5))
    host = torch.empty((ROWS_PER_GPU, N), dtype=torch.complex64, pin_memory=True)
    host.copy_(x)
    h = host.numpy()
    fft.fft_batch(h)  # warm-up: allocates the staging workspace
    host.copy_(x)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fft.fft_batch(h)  # H2D 2 GiB -> kernel -> D2H 2 GiB, synchronous
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    nbytes = ROWS_PER_GPU * N * 8
    e2e = {"value": FLOPS_PER_TRANSFORM * ROWS_PER_GPU * n_gpus / e2e_s / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_s * 1e3,
           "steps": e2e_steps, "api": "kofft_cuda_fft_batch_host_f32 (CudaFftImpl.fft_batch on pinned host rows; "
                                      "32 MiB chunks pipelined over H2D / kernel / D2H streams)"}
    del host, h

    # ---- extra: STFT (configs[3] shape) and FAST-mode C2C, after the headline region ------------
    extra = {}
    try:
        fast = kofft_b200.CudaFftImpl(device=local, exact=False)
        for _ in range(3):
            fast.fft_batch(x, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fast.fft_batch(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        extra["c2c_fast_mode"] = {"ms_per_step": ms, "gflops": FLOPS_PER_TRANSFORM * ROWS_PER_GPU / ms / 1e6,
                                  "hbm_gbs": ALGO_BYTES_PER_TRANSFORM * ROWS_PER_GPU / ms / 1e6,
                                  "frac_of_measured_peak": ALGO_BYTES_PER_TRANSFORM * ROWS_PER_GPU / ms / 1e6 / peak}
        del x, y
        torch.cuda.empty_cache()
        # f64 twin: C2C N = 4096 x 16384 complex128 rows (1 GiB in + 1 GiB out)
        fft64 = kofft_b200.CudaFftImpl64(ctx=fft.ctx)
        dn, db = 4096, 16384
        xd = torch.view_as_complex((torch.rand((db, dn, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1).contiguous())
        yd = torch.empty_like(xd)
        for _ in range(3):
            fft64.fft_batch(xd, out=yd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fft64.fft_batch(xd, out=yd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        extra["c2c_f64_4096x16384"] = {"ms": ms, "gflops": 5 * dn * 12 * db / ms / 1e6, "hbm_gbs": 32 * dn * db / ms / 1e6,
                                       "frac_of_measured_peak": 32 * dn * db / ms / 1e6 / peak,
                                       "note": "f64 twin of the single-CTA engine (fft_f64.cuh), bit-identical to the f64 oracle"}
        del xd, yd
        torch.cuda.empty_cache()
        # rfft N = 2^16 x batch 16384 (BASELINE configs[2]): two-pass path with the fused twist
        rn, rb = 65536, 16384
        xr = (torch.rand((rb, rn), generator=g, device=dev) * 2 - 1).contiguous()
        yr = torch.empty((rb, rn // 2 + 1), dtype=torch.complex64, device=dev)
        for _ in range(2):
            fft.rfft_batch(xr, out=yr)
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fft.rfft_batch(xr, out=yr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        algo = (4 * rn + 8 * (rn // 2 + 1)) * rb
        extra["rfft_65536x16384"] = {"ms": ms, "gflops_nominal": 2.5 * rn * 16 * rb / ms / 1e6, "hbm_gbs": algo / ms / 1e6,
                                     "frac_of_measured_peak": algo / ms / 1e6 / peak,
                                     "note": "default path: one persistent cooperative kernel, column pass of chunk p "
                                             "overlapped with row pass + fused Hermitian twist of chunk p-1, "
                                             "intermediate pinned in L2"}
        for key, mode_id, note in (
                ("rfft_65536x16384_two_kernel_path", 0, "column pass + row pass, two kernels per 256 MB batch chunk"),
                ("rfft_65536x16384_cluster_kernel_path", 1,
                 "one persistent thread-block-cluster kernel (column pass, cluster barrier, row pass + twist)")):
            fft.ctx.set_large_mode(mode_id)
            fft.rfft_batch(xr, out=yr)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fft.rfft_batch(xr, out=yr)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / reps
            extra[key] = {"ms": ms2, "hbm_gbs": algo / ms2 / 1e6, "frac_of_measured_peak": algo / ms2 / 1e6 / peak, "note": note}
        fft.ctx.set_large_mode(3)
        del xr, yr
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info()
        ch, length, hop, win = 64, 28_800_000, 512, 2048
        if free < 90e9:
            ch = 8
        nframes = -(-length // hop)
        sig = (torch.rand((ch, length), generator=g, device=dev) * 2 - 1).contiguous()
        w = torch.from_numpy(W.hann(win)).to(dev)
        frames = torch.empty((ch, nframes, win), dtype=torch.complex64, device=dev)
        S.stft_batch(fft, sig, w, hop, nframes, out=frames)
        torch.cuda.synchronize()
        reps = 6
        stft_sampler = ClockSampler(local)  # the STFT is FP32-bound: its time moves with the SM clock
        stft_sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            S.stft_batch(fft, sig, w, hop, nframes, out=frames)
        e1.record()
        torch.cuda.synchronize()
        stft_clocks = stft_sampler.stop()
        ms = e0.elapsed_time(e1) / reps
        algo = 4 * ch * length + 8 * ch * nframes * win
        # FP32 floor of the EXACT kernel: 10 lane-operations per butterfly, (N/2) log2 N butterflies per
        # frame minus the real-input shortcuts (8 %), 128 lanes per SM and clock
        fma_floor_ms = {mhz: 10 * 0.92 * (win // 2) * 11 * ch * nframes / (128 * 148 * mhz * 1e6) * 1e3
                        for mhz in (stft_clocks.get("sm_mhz") or 1965, 1965)}
        cpu_stft = {}
        if rank == 0:
            try:
                threads_c = os.cpu_count() or 1
                fps_shared, smp = cpu_port_stft_frames_per_s(threads_c, False)
                fps_fresh, _ = cpu_port_stft_frames_per_s(threads_c, True)
                cpu_stft = {"frames_per_s_shared_table": fps_shared, "frames_per_s_fresh_planner_per_frame": fps_fresh,
                            "cores": threads_c, "kind": "port", "sample": smp}
            except Exception as e:  # pragma: no cover
                cpu_stft = {"error": repr(e)}
        extra["stft"] = {"workload": f"Hann {win}, hop {hop}, {ch} ch x {length} samples (BASELINE configs[3])",
                         "cpu_baseline": cpu_stft,
                         "frames_per_s": ch * nframes / (ms * 1e-3), "ms": ms, "hbm_gbs": algo / ms / 1e6,
                         "frac_of_measured_peak": algo / ms / 1e6 / peak, "clocks": stft_clocks,
                         "fp32_floor_ms_at_measured_clock": fma_floor_ms[stft_clocks.get("sm_mhz") or 1965],
                         "fp32_floor_ms_at_max_clock": fma_floor_ms[1965],
                         "hbm_floor_ms": algo / peak / 1e6}
        S.stft_batch(fast, sig, w, hop, nframes, out=frames)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            S.stft_batch(fast, sig, w, hop, nframes, out=frames)
        e1.record()
        torch.cuda.synchronize()
        msf = e0.elapsed_time(e1) / reps
        extra["stft_fast_mode"] = {"frames_per_s": ch * nframes / (msf * 1e-3), "ms": msf, "hbm_gbs": algo / msf / 1e6,
                                   "frac_of_measured_peak": algo / msf / 1e6 / peak}
        out = torch.zeros((ch, length), device=dev)
        S.istft_batch(fft, frames, w, hop, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            S.istft_batch(fft, frames, w, hop, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        extra["istft"] = {"frames_per_s": ch * nframes / (ms * 1e-3), "ms": ms, "hbm_gbs": algo / ms / 1e6,
                          "frac_of_measured_peak": algo / ms / 1e6 / peak,
                          "note": "one fused kernel (ifft + window + ordered overlap-add + normalisation)"}
        fft.ctx.set_istft_fusion(False)
        S.istft_batch(fft, frames, w, hop, out)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            S.istft_batch(fft, frames, w, hop, out)
        e1.record()
        torch.cuda.synchronize()
        fft.ctx.set_istft_fusion(True)
        ms2 = e0.elapsed_time(e1) / reps
        extra["istft_two_kernel_path"] = {"frames_per_s": ch * nframes / (ms2 * 1e-3), "ms": ms2,
                                          "frac_of_measured_peak": algo / ms2 / 1e6 / peak}
        del sig, frames, out
    except Exception as e:  # extras never invalidate the headline line
        extra["error"] = f"{type(e).__name__}: {e}"

    # ---- extra (N > 1): ONE transform sharded over the ranks (BASELINE configs[4] shape: 2^27 points per GPU) --
    if world > 1 and (world & (world - 1)) == 0:
        try:
            from kofft_b200 import dist as KD

            torch.cuda.empty_cache()
            log2n = 27 + int(math.log2(world))
            dfft = KD.DistFft(fft.ctx, rank, world, log2n)
            dfft.connect()
            shard = (1 << log2n) // world
            xs = torch.view_as_complex(torch.rand((shard, 2), generator=g, device=dev) * 2 - 1).contiguous()
            os_ = torch.empty_like(xs)
            walls = []
            for it in range(5):
                barrier()
                t0 = time.perf_counter()
                dfft.transform(xs, out=os_, natural_order=True)
                if it >= 2:
                    walls.append(time.perf_counter() - t0)
            w = max_over_ranks(statistics.median(walls))
            extra["dist_c2c_one_transform"] = {
                "log2n": log2n, "points_per_gpu": shard, "ms": w * 1e3, "gflops": 5.0 * (1 << log2n) * log2n / w / 1e9,
                "note": "four-step split; all-to-all exchanges are P2P stores issued by the transpose-scatter kernels "
                        "(CUDA IPC over NVLink), host barriers between phases; natural-order output (3 exchanges); "
                        "validated against f64, no kofft reference at this size (SURVEY 0.5)"}
            dfft.close()
        except Exception as e:
            extra["dist_error"] = f"{type(e).__name__}: {e}"

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ------------------------
    cpu = None
    if rank == 0 and n_gpus == 1:
        threads = os.cpu_count() or 1
        rows = 8192
        reps = 1
        gf, dt = cpu_port_gflops(rows, reps, threads)
        while dt < 10.0 and reps < 4096:  # aim at >= 10 s of CPU work in the timed sample
            reps *= 2
            gf, dt = cpu_port_gflops(rows, reps, threads)
        cpu = {"value": gf, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{rows} rows x {reps} reps of N={N} ({dt:.1f} s), {threads} threads over rows, one planner "
                         "per thread; oracle/kofft_oracle.c (C restatement; the Rust crate cannot be built here)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n_gpus),
            "hbm_gbs": ALGO_BYTES_PER_TRANSFORM * ROWS_PER_GPU * n_gpus / (ms_per_step * 1e-3) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: anything a library writes to file descriptor 1 while
    # the benchmark runs (NCCL prints its version banner there) is sent to stderr instead
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    out = buf.getvalue()
    if out:
        sys.stdout.write(out if out.endswith("\n") else out + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
