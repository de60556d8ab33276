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


HEADLINE_KERNEL = "fft_cta_kernelILi12ELb1ENS_5IoC2CILb0EEELb1EE"  # mangled-name fragment of the headline instantiation


def headline_sass_hash() -> str | None:
    """sha256 of the headline kernel's SASS in the library that is loaded: the committed ncu traffic figure is only
    quoted while the kernel it was measured on is the kernel that runs (scripts/summarize_ncu.py stamps it)."""
    import hashlib
    import subprocess

    lib = os.environ.get("KOFFT_CUDA_LIB") or os.path.join(ROOT, "kofft_b200", "lib", "libkofft_cuda.so")
    try:
        names = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True, timeout=120).stdout
        fn = next((w for w in names.split() if HEADLINE_KERNEL in w and w.startswith(".text.")), None)
        if fn is None:
            return None
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn[len(".text."):], lib], capture_output=True, text=True,
                              timeout=300).stdout
        body = "\n".join(l.split("/*")[1].split("*/")[1].strip() if l.count("/*") >= 2 else "" for l in sass.splitlines()
                         if "/*" in l and "Function" not in l)
        return hashlib.sha256(body.encode()).hexdigest() if body.strip() else None
    except Exception:
        return None


def ncu_traffic() -> tuple[float | None, str]:
    """dram bytes per launch of the headline kernel from the committed ncu summary -- only if that capture was
    taken on the same SASS as the kernel that runs now; otherwise None (stale)."""
    path = os.path.join(ROOT, "profiles", "headline_kernel_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        want = d.get("sass_sha256")
        if not want:
            return None, "committed ncu capture carries no SASS stamp"
        have = headline_sass_hash()
        if have != want:
            return None, "stale: the committed ncu capture was taken on different SASS of the headline kernel"
        return float(d["dram_bytes_per_launch"]), "ncu --set full capture of this SASS (profiles/%s)" % d.get("source", "")
    except Exception as e:
        return None, f"no committed capture ({type(e).__name__})"


def bind_to_gpu_numa_node(index: int, props=None) -> dict:
    """Pin this rank (and with it the pages of the pinned buffers it allocates afterwards: first touch) to the NUMA
    node of its GPU.  Without this every rank of an 8-GPU run allocates on node 0 and the host side of the e2e
    number collapses (SCALE_r01: 257 ms/step at 8 GPUs against 45 at 1)."""
    info = {"node": None, "cpus_bound": False, "mem_bound": False}
    try:
        import ctypes

        if props is not None and hasattr(props, "pci_bus_id"):  # the CUDA device's own PCI address (index != NVML index
            bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"  # under CUDA_VISIBLE_DEVICES)
        else:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bus = bus.lower()
            if len(bus.split(":")[0]) == 8:
                bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        info["node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus_bound"] = True
            info["cpus"] = len(allowed)
        # set_mempolicy(MPOL_PREFERRED, {node}): pages of later allocations come from the GPU's node
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))  # __NR_set_mempolicy (x86-64), MPOL_PREFERRED
        info["mem_bound"] = rc == 0
    except Exception as e:
        info["error"] = f"{type(e).__name__}: {e}"
    return info


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
_CPU_ROWS: dict = {}


def _cpu_rows(rows: int) -> np.ndarray:
    """uniform[-1,1) complex64 rows for the CPU legs, generated once (2 GiB for the full step)"""
    if rows not in _CPU_ROWS:
        rng = np.random.default_rng(0)
        a = rng.random((rows, 2 * N), dtype=np.float32)
        a *= 2.0
        a -= 1.0
        _CPU_ROWS.clear()
        _CPU_ROWS[rows] = a.view(np.complex64)
    return _CPU_ROWS[rows]


def cpu_port_gflops(rows: int, reps: int, threads: int) -> tuple[float, float]:
    """kofft's CPU path (oracle port) on `rows` transforms of length N, all `threads` cores:
    rows split evenly over threads, one reused planner per thread.  Returns (GFLOP/s, seconds)."""
    from oracle import kofft_oracle as ko

    ko.build()
    x = _cpu_rows(rows)
    # fast=True: the -O3 / 128-bit-vector build of the same sources (the reference's hot loop is explicit SSE,
    # src/fft.rs:845-862); bit-identical to the -O2 oracle (tests/test_oracle_golden.py)
    ko.fft_batch_inplace(x[: max(threads, 1)].copy(), nthreads=threads, fast=True)  # warm the threads / tables
    t0 = time.perf_counter()
    for _ in range(reps):
        ko.fft_batch_inplace(x, nthreads=threads, fast=True)
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
    ko.stft_batch(sig[:, :48_000], w, hop, -(-48_000 // hop), fresh_planner=fresh_planner, nthreads=threads, fast=True)  # warm-up
    t0 = time.perf_counter()
    ko.stft_batch(sig, w, hop, nframes, fresh_planner=fresh_planner, nthreads=threads, fast=True)
    dt = time.perf_counter() - t0
    return ch * nframes / dt, f"{ch} ch x {length} samples ({ch * nframes} frames, {dt:.1f} s), {threads} threads over channels"


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = ROWS_PER_GPU  # the whole step: 65536 rows take a fraction of a second on the host cores
    for _ in range(max(args.warmup, 1)):
        cpu_port_gflops(rows, 1, threads)
    t0 = time.perf_counter()
    vals = [cpu_port_gflops(rows, 1, threads)[0] for _ in range(args.steps)]
    total = time.perf_counter() - t0
    value = statistics.median(vals)
    sample = (f"all {rows} rows of the step (N={N}), {threads} threads, rows split evenly, one planner per thread; "
              "-O3 / 128-bit-vector build of the port, bit-identical to the oracle")
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
                "bit-faithful C restatement of its CPU path (oracle/kofft_oracle.c, built -O3 with 128-bit vectors like the "
                "reference's explicit SSE loop) on the host cores.  kofft itself has no parallel batched FFT "
                "(batch() is sequential, src/fft.rs:2156-2164): the thread pool over rows is the generous reading",
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
    orig_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local, torch.cuda.get_device_properties(local))
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
    traffic, traffic_src = ncu_traffic() if rank == 0 else (None, "")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
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
    # the ceiling of this API shape: the same bytes as plain pinned copies, H2D and D2H at the same time on two
    # streams, every rank at once -- what the host side of the box can deliver to N GPUs
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    host2 = torch.empty((ROWS_PER_GPU, N), dtype=torch.complex64, pin_memory=True)
    dbuf = torch.empty((ROWS_PER_GPU, N), dtype=torch.complex64, device=dev)

    dbuf2 = torch.empty((ROWS_PER_GPU, N), dtype=torch.complex64, device=dev)

    def copies():
        with torch.cuda.stream(s_in):
            dbuf.copy_(host, non_blocking=True)
        with torch.cuda.stream(s_out):
            host2.copy_(dbuf2, non_blocking=True)

    copies()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copies()
        torch.cuda.synchronize()
    barrier()
    ceil_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e["pcie_ceiling_ms_per_step"] = ceil_s * 1e3
    e2e["pcie_ceiling_value"] = FLOPS_PER_TRANSFORM * ROWS_PER_GPU * n_gpus / ceil_s / 1e9
    e2e["frac_of_pcie_ceiling"] = ceil_s / e2e_s
    e2e["pcie_gbs_per_rank_each_way"] = nbytes / e2e_s / 1e9
    e2e["pcie_ceiling_gbs_per_rank_each_way"] = nbytes / ceil_s / 1e9
    e2e["numa"] = numa
    del host, h, host2, dbuf, dbuf2

    # ---- extra: STFT (configs[3] shape) and FAST-mode C2C, after the headline region ------------
    extra = {}
    try:
        if world > 1:
            # strong scaling of configs[1]: the SAME 65536 rows cut over the ranks (65536 / N rows each)
            rows_s = ROWS_PER_GPU // world
            xs_, ys_ = x[:rows_s], y[:rows_s]
            for _ in range(3):
                fft.fft_batch(xs_, out=ys_)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                fft.fft_batch(xs_, out=ys_)
            e1.record()
            barrier()
            ms_s = max_over_ranks(e0.elapsed_time(e1) / args.steps)
            extra["c2c_strong_scaling"] = {
                "workload": f"N={N} x batch {ROWS_PER_GPU} in total, {rows_s} rows per GPU (BASELINE configs[1] cut over the ranks)",
                "ms_per_step": ms_s, "gflops": FLOPS_PER_TRANSFORM * ROWS_PER_GPU / ms_s / 1e6,
                "speedup_vs_one_gpu_same_run": ms_per_step / ms_s, "efficiency": ms_per_step / ms_s / world,
                "hbm_frac_per_gpu": ALGO_BYTES_PER_TRANSFORM * rows_s / ms_s / 1e6 / peak,
                "note": "one-GPU time = this run's 65536-row step on every rank (max over ranks); no collective"}
        if rank == 0:
            # BASELINE configs[0] (examples/basic_usage.rs:232-238): one 1024-point FFT + IFFT through the trait-level
            # host-pointer call; latency is launch + two small PCIe copies + synchronisation
            xin = (np.sin(0.1 * np.arange(1024, dtype=np.float32)) + 0j).astype(np.complex64)
            buf = xin.copy()
            for _ in range(20):
                fft.fft(buf)
                fft.ifft(buf)
            lat = []
            for _ in range(200):
                buf[:] = xin
                t0 = time.perf_counter()
                fft.fft(buf)
                fft.ifft(buf)
                lat.append((time.perf_counter() - t0) * 1e6)
            lat.sort()
            dx = torch.from_numpy(xin).to(dev).reshape(1, 1024)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(10):
                fft.fft_batch(dx)
            e0.record()
            for _ in range(100):
                fft.fft_batch(dx)
                fft.fft_batch(dx, inverse=True)
            e1.record()
            torch.cuda.synchronize()
            extra["config1_latency"] = {
                "workload": "ScalarFftImpl-style 1024-point C2C FFT then IFFT, batch 1 (BASELINE configs[0])",
                "host_call_us_median": lat[len(lat) // 2], "host_call_us_p90": lat[int(len(lat) * 0.9)],
                "host_call_us_min": lat[0], "device_resident_us_per_fft_ifft_pair": e0.elapsed_time(e1) * 1e3 / 100,
                "api": "kofft_cuda_fft_host_f32 x 2 (host buffer in, host buffer out, synchronous)",
                "max_roundtrip_err": float(np.abs(buf - xin).max())}
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
                                     "note": "default path (fft_split32.cuh): one persistent warp-specialised kernel -- A warps run "
                                             "the 1024-point column transforms of TMA-loaded tiles, B warps the 32-point rows in "
                                             "registers with the Hermitian twist by warp shuffles; intermediate kept in L2",
                                     "fallbacks": fft.ctx.fallback_count}
        yi = torch.empty_like(xr)
        for _ in range(2):
            fft.irfft_batch(yr, rn, out=yi)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fft.irfft_batch(yr, rn, out=yi)
        e1.record()
        torch.cuda.synchronize()
        msi = e0.elapsed_time(e1) / reps
        extra["irfft_65536x16384"] = {"ms": msi, "hbm_gbs": algo / msi / 1e6, "frac_of_measured_peak": algo / msi / 1e6 / peak,
                                      "note": "split kernel (fft_split32.cuh): its B warps untwist the rows three transforms ahead of "
                                              "pass A into L2-resident rows that pass A stages by tensor-map copies (older "
                                              "pipelined path: 4.4 ms)"}
        del yi
        for key, min_l, mode_id, note in (
                ("rfft_65536x16384_pipelined_16_per_thread_path", 16, 2,
                 "round-1 default: persistent cooperative kernel, 16 elements per thread, teams of 8 CTAs"),
                ("rfft_65536x16384_two_kernel_path", 16, 0, "column pass + row pass, two kernels per 256 MB batch chunk")):
            fft.ctx.set_split_min_log2n(min_l)
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
        fft.ctx.set_split_min_log2n(14)
        del xr, yr
        # C2C at the sizes between the single-CTA kernel and the headline config, and one length above 2^16
        for cn in (8192, 16384, 32768, 65536, 1 << 20):
            crows = (1 << 28) // cn
            xc = torch.view_as_complex(torch.rand((crows, cn, 2), generator=g, device=dev) * 2 - 1).contiguous()
            yc = torch.empty_like(xc)
            for _ in range(2):
                fft.fft_batch(xc, out=yc)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fft.fft_batch(xc, out=yc)
            e1.record()
            torch.cuda.synchronize()
            msc = e0.elapsed_time(e1) / reps
            kernel_of = {8192: "fft_wide_kernel (32 elements per thread, one CTA per transform, two CTAs per SM)",
                         16384: "fft_wide_kernel (32 elements per thread, one 512-thread CTA per transform)",
                         32768: "split32_kernel (warp-specialised, teams of 4 CTAs)",
                         65536: "colpass_kernel + rowpass_kernel", 1 << 20: "colpass_kernel + 2 x huge_pass_kernel"}
            extra[f"c2c_{cn}x{crows}"] = {"ms": msc, "kernel": kernel_of[cn], "gflops": 5.0 * cn * math.log2(cn) * crows / msc / 1e6,
                                          "hbm_gbs": 16.0 * cn * crows / msc / 1e6,
                                          "frac_of_measured_peak": 16.0 * cn * crows / msc / 1e6 / peak}
            del xc, yc
        # rfft / irfft at 2^14 reals (wide single-CTA kernel: twist through a half-row side buffer, untwist from staged bins)
        mrn, mrb = 16384, 32768
        xm = (torch.rand((mrb, mrn), generator=g, device=dev) * 2 - 1).contiguous()
        ym = torch.empty((mrb, mrn // 2 + 1), dtype=torch.complex64, device=dev)
        for key, fn in (("rfft_16384x32768", lambda: fft.rfft_batch(xm, out=ym)), ("irfft_16384x32768", lambda: fft.irfft_batch(ym, mrn, out=xm))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            msm = e0.elapsed_time(e1) / reps
            algom = (4 * mrn + 8 * (mrn // 2 + 1)) * mrb
            extra[key] = {"ms": msm, "kernel": "fft_wide_kernel", "hbm_gbs": algom / msm / 1e6, "frac_of_measured_peak": algom / msm / 1e6 / peak}
        del xm, ym
        if rank == 0:  # the reference's own benchmark size N = 2^20 (benchmarks/README.md:5-7) beside the CPU port
            try:
                from oracle import kofft_oracle as ko

                one = (np.random.default_rng(20).uniform(-1, 1, (4, 1 << 20)) + 0j).astype(np.complex64)
                ko.fft_batch_inplace(one.copy(), nthreads=1, fast=True)
                t0 = time.perf_counter()
                ko.fft_batch_inplace(one, nthreads=1, fast=True)
                dtc = (time.perf_counter() - t0) / 4
                extra["c2c_1048576x256"]["cpu_port_one_thread_ms_per_transform"] = dtc * 1e3
                extra["c2c_1048576x256"]["gpu_ms_per_transform"] = extra["c2c_1048576x256"]["ms"] / 256
            except Exception as e:  # pragma: no cover
                extra["c2c_1048576x256"]["cpu_error"] = repr(e)
        if rank == 0:
            # ONE transform of 2^20 points, the shape of the reference's own published table (C2C 59.3 ms,
            # real 66.9 ms per transform on its authors' CPU): device-resident row, and host row in / host row out
            one_c = torch.view_as_complex(torch.rand((1, 1 << 20, 2), generator=g, device=dev) * 2 - 1).contiguous()
            one_o = torch.empty_like(one_c)
            one_r = (torch.rand((1, 1 << 20), generator=g, device=dev) * 2 - 1).contiguous()
            lat = {}
            for key, fn in (("c2c_device_us", lambda: fft.fft_batch(one_c, out=one_o)), ("rfft_device_us", lambda: fft.rfft_batch(one_r))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(20):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                lat[key] = e0.elapsed_time(e1) / 20 * 1e3
            hrow = (np.random.default_rng(21).uniform(-1, 1, (1, 1 << 20)) + 0j).astype(np.complex64)
            fft.fft_batch(hrow.copy())
            ts = []
            for _ in range(5):
                hh = hrow.copy()
                t0 = time.perf_counter()
                fft.fft_batch(hh)
                ts.append(time.perf_counter() - t0)
            lat["c2c_host_call_us_median"] = sorted(ts)[len(ts) // 2] * 1e6
            lat["published_reference_ms"] = {"c2c": 59.3, "real": 66.9, "source": "BASELINE.md (benchmarks/README.md:67-70), other hardware"}
            extra["single_transform_1048576"] = lat
            del one_c, one_o, one_r
        torch.cuda.empty_cache()

        # ---- STFT / ISTFT (BASELINE configs[3]); N > 1: the 64 channels are sharded over the ranks ----------------
        free, _ = torch.cuda.mem_get_info()
        ch_total, length, hop, win = 64, 28_800_000, 512, 2048
        ch = ch_total // world if (world > 1 and ch_total % world == 0) else ch_total
        if free < 90e9 and world == 1:
            ch = ch_total = 8
        nframes = -(-length // hop)
        sig = (torch.rand((ch, length), generator=g, device=dev) * 2 - 1).contiguous()
        w = torch.from_numpy(W.hann(win)).to(dev)
        frames = torch.empty((ch, nframes, win), dtype=torch.complex64, device=dev)
        S.stft_batch(fft, sig, w, hop, nframes, out=frames)
        barrier()
        reps = 6
        stft_sampler = ClockSampler(local)  # the STFT is FP32-bound: its time moves with the SM clock
        stft_sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            S.stft_batch(fft, sig, w, hop, nframes, out=frames)
        e1.record()
        barrier()
        stft_clocks = stft_sampler.stop()
        ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        algo = 4 * ch * length + 8 * ch * nframes * win  # per GPU
        # FP32 floor of the EXACT kernel: 10 lane-operations per butterfly, (N/2) log2 N butterflies per
        # frame minus the real-input shortcuts (8 %), 128 lanes per SM and clock
        fma_floor_ms = {mhz: 10 * 0.92 * (win // 2) * 11 * ch * nframes / (128 * 148 * mhz * 1e6) * 1e3
                        for mhz in (stft_clocks.get("sm_mhz") or 1965, 1965)}
        cpu_stft = {}
        if rank == 0 and world == 1:
            try:
                threads_c = os.cpu_count() or 1
                os.sched_setaffinity(0, orig_affinity)  # the CPU legs use every host core, not just the GPU's NUMA node
                fps_shared, smp = cpu_port_stft_frames_per_s(threads_c, False)
                fps_fresh, _ = cpu_port_stft_frames_per_s(threads_c, True)
                cpu_stft = {"frames_per_s_shared_table": fps_shared, "frames_per_s_fresh_planner_per_frame": fps_fresh,
                            "cores": threads_c, "kind": "port", "sample": smp}
            except Exception as e:  # pragma: no cover
                cpu_stft = {"error": repr(e)}
        shard_note = (f"{ch_total} channels sharded over {world} GPUs ({ch} each), no collective; time = max over ranks"
                      if world > 1 else f"{ch} channels on one GPU")
        extra["stft"] = {"workload": f"Hann {win}, hop {hop}, {ch_total} ch x {length} samples (BASELINE configs[3]); " + shard_note,
                         "mode": "exact (bit-identical to the reference)", "cpu_baseline": cpu_stft,
                         "frames_per_s": ch * world * nframes / (ms * 1e-3), "ms": ms, "hbm_gbs_per_gpu": algo / ms / 1e6,
                         "frac_of_measured_peak": algo / ms / 1e6 / peak, "clocks": stft_clocks,
                         "fp32_floor_ms_at_measured_clock": fma_floor_ms[stft_clocks.get("sm_mhz") or 1965],
                         "fp32_floor_ms_at_max_clock": fma_floor_ms[1965],
                         "hbm_floor_ms": algo / peak / 1e6}
        S.stft_batch(fast, sig, w, hop, nframes, out=frames)
        barrier()
        e0.record()
        for _ in range(reps):
            S.stft_batch(fast, sig, w, hop, nframes, out=frames)
        e1.record()
        barrier()
        msf = max_over_ranks(e0.elapsed_time(e1) / reps)
        # parity of the FAST mode against EXACT (which is bit-identical to the oracle): whole frames of a channel subset
        pc = min(ch, 4)
        fe = S.stft_batch(fft, sig[:pc], w, hop, nframes)
        ff = frames[:pc]
        d2 = (torch.view_as_real(ff) - torch.view_as_real(fe)).double().pow(2).sum(dim=(2, 3))
        r2 = torch.view_as_real(fe).double().pow(2).sum(dim=(2, 3))
        stft_fast_parity = {"rel_l2_vs_exact": float(torch.sqrt(d2.sum() / r2.sum()).item()),
                            "worst_frame_rel_l2_vs_exact": float(torch.sqrt((d2 / r2.clamp_min(1e-30)).max()).item()),
                            "frames_compared": int(pc * nframes), "tolerance": 1e-5}
        del fe, ff, d2, r2
        extra["stft_fast_mode"] = {"mode": "fast (fused multiply-add butterflies; inside the north star's 1e-5 rel-L2)",
                                   "frames_per_s": ch * world * nframes / (msf * 1e-3), "ms": msf, "hbm_gbs_per_gpu": algo / msf / 1e6,
                                   "frac_of_measured_peak": algo / msf / 1e6 / peak, "parity": stft_fast_parity}
        S.stft_batch(fft, sig, w, hop, nframes, out=frames)  # exact frames for the inverse
        out = torch.zeros((ch, length), device=dev)
        S.istft_batch(fft, frames, w, hop, out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            S.istft_batch(fft, frames, w, hop, out)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        out.zero_()
        S.istft_batch(fft, frames, w, hop, out)
        inner = slice(win, length - win)
        rt = float(((out[:, inner] - sig[:, inner]).double().norm() / sig[:, inner].double().norm()).item())
        extra["istft"] = {"mode": "exact", "frames_per_s": ch * world * nframes / (ms * 1e-3), "ms": ms, "hbm_gbs_per_gpu": algo / ms / 1e6,
                          "frac_of_measured_peak": algo / ms / 1e6 / peak, "roundtrip_rel_l2_vs_signal": rt,
                          "note": "one fused kernel (ifft + window + ordered overlap-add + normalisation); " + shard_note}
        exact_out = out[:min(ch, 4)].clone()
        out.zero_()
        S.istft_batch(fast, frames, w, hop, out)
        barrier()
        e0.record()
        for _ in range(reps):
            S.istft_batch(fast, frames, w, hop, out)
        e1.record()
        barrier()
        msf = max_over_ranks(e0.elapsed_time(e1) / reps)
        out.zero_()
        S.istft_batch(fast, frames, w, hop, out)
        pf = float(((out[:exact_out.shape[0]] - exact_out).double().norm() / exact_out.double().norm()).item())
        extra["istft_fast_mode"] = {"mode": "fast", "frames_per_s": ch * world * nframes / (msf * 1e-3), "ms": msf,
                                    "frac_of_measured_peak": algo / msf / 1e6 / peak,
                                    "parity": {"rel_l2_vs_exact": pf, "tolerance": 1e-5}}
        del exact_out
        if world == 1:
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
        torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            # STFT end to end: pinned host signal -> pinned host frames through kofft_cuda_stft_host_f32 (one copy in,
            # one kernel, one copy out); a bounded sample of configs[3] (4 channels x 10 min: 0.46 GB in, 3.7 GB out)
            ech = 4
            hs = torch.empty((ech, length), dtype=torch.float32, pin_memory=True)
            hs.uniform_(-1, 1)
            hf = torch.empty((ech, nframes, win), dtype=torch.complex64, pin_memory=True)
            hw = W.hann(win)
            S.stft_batch(fft, hs.numpy(), hw, hop, nframes, out=hf.numpy())
            t0 = time.perf_counter()
            for _ in range(3):
                S.stft_batch(fft, hs.numpy(), hw, hop, nframes, out=hf.numpy())
            dte = (time.perf_counter() - t0) / 3
            extra["stft_e2e"] = {"frames_per_s": ech * nframes / dte, "ms": dte * 1e3, "h2d_bytes": ech * length * 4,
                                 "d2h_bytes": ech * nframes * win * 8,
                                 "sample": f"{ech} of 64 channels x {length} samples, pinned host buffers",
                                 "api": "kofft_cuda_stft_host_f32 (H2D signal, fused framing + window + FFT kernel, D2H frames)",
                                 "pcie_gbs_d2h": ech * nframes * win * 8 / dte / 1e9}
            del hs, hf
    except Exception as e:  # extras never invalidate the headline line
        extra["error"] = f"{type(e).__name__}: {e}"

    # ---- extra (N > 1): ONE transform sharded over the ranks (BASELINE configs[4] shape: 2^27 points per GPU) --
    if world > 1 and (world & (world - 1)) == 0:
        try:
            from kofft_b200 import dist as KD
            from kofft_b200 import dist_validate

            torch.cuda.empty_cache()
            log2n = 27 + int(math.log2(world))
            dfft = KD.DistFft(fft.ctx, rank, world, log2n)
            dfft.connect()
            shard = (1 << log2n) // world
            xs = torch.view_as_complex(torch.rand((shard, 2), generator=g, device=dev) * 2 - 1).contiguous()

            def timed(fn, iters=5, skip=2):
                walls = []
                for it in range(iters):
                    barrier()
                    t0 = time.perf_counter()
                    fn()
                    torch.cuda.synchronize()
                    if it >= skip:
                        walls.append(time.perf_counter() - t0)
                return max_over_ranks(statistics.median(walls))

            t_p2p = timed(lambda: dfft.transform(xs, natural_order=True))
            t_p2p_t = timed(lambda: dfft.transform(xs, natural_order=False))
            t_col = timed(lambda: dfft.transform_collective(xs))
            same = torch.equal(torch.view_as_real(dfft.transform_collective(xs).clone()),
                               torch.view_as_real(dfft.transform(xs, natural_order=True)))
            link_bytes = shard * 8 * (world - 1) // world  # per GPU, per direction, per exchange
            check = dist_validate.validate(dfft)           # closed forms + f64 direct sums at the full size
            extra["dist_c2c_one_transform"] = {
                "log2n": log2n, "points_per_gpu": shard, "ms": t_p2p * 1e3,
                "gflops": 5.0 * (1 << log2n) * log2n / t_p2p / 1e9,
                "ms_transposed_order_output": t_p2p_t * 1e3,
                "ms_nccl_all_to_all_arm": t_col * 1e3,
                "p2p_store_speedup_over_nccl_arm": t_col / t_p2p,
                "arms_bit_identical": bool(same),
                "nvlink_bytes_per_gpu_per_exchange": link_bytes,
                "nvlink_floor_ms_natural_order_at_770GBs": 3 * link_bytes / 770e9 * 1e3,
                "validation": check,
                "note": "four-step split, natural-order output (3 exchanges).  'ms': the exchanges are peer-to-peer stores issued by "
                        "the transpose-scatter kernels (CUDA IPC over NVLink) overlapped with the local transforms; "
                        "'ms_nccl_all_to_all_arm': the same phases with pack -> torch.distributed.all_to_all_single (ncclAlltoAll) "
                        "-> unpack.  Validation at the FULL size: K tones (exact spikes), an impulse (phase ramp at every bin), "
                        "uniform noise with 64 bins against f64 direct sums over all N points; no kofft reference exists at "
                        "this size (SURVEY 0.5)"}
            dfft.close()
        except Exception as e:
            extra["dist_error"] = f"{type(e).__name__}: {e}"

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ------------------------
    cpu = None
    if rank == 0 and n_gpus == 1:
        os.sched_setaffinity(0, orig_affinity)  # every host core, not just the GPU's NUMA node
        threads = os.cpu_count() or 1
        rows = ROWS_PER_GPU  # the whole step
        reps = 1
        gf, dt = cpu_port_gflops(rows, reps, threads)
        while dt < 10.0 and reps < 4096:  # aim at >= 10 s of CPU work in the timed sample
            reps *= 2
            gf, dt = cpu_port_gflops(rows, reps, threads)
        cpu = {"value": gf, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"all {rows} rows of the step x {reps} reps of N={N} ({dt:.1f} s), {threads} threads over rows, one planner "
                         "per thread; oracle/kofft_oracle.c built -O3 with 128-bit vectors (bit-identical to the -O2 oracle; "
                         "the Rust crate cannot be built here)"}

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
