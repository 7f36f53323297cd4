#!/usr/bin/env python
"""bench.py -- headline benchmark of the Diagonalize hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   # the reference on host cores

Workload (config.workload): 16M 4x4 fp64 quaternion-superposition overlap matrices PER GPU
(weak scaling: rank r diagonalises slice r of an N*16M-matrix counter-generated stream, no
inter-GPU traffic), eigenvectors on, SORT_DECREASING_EVALS, max_num_sweeps=50.
One "step" = one pass of the hot path over the rank's whole batch.

JSON line (rank 0): value = whole-job matrices/s with inputs resident in HBM (SoA-interleaved
device layout, one kernel launch per step, CUDA events, max over ranks);  e2e = the same
metric through the host-buffer C-ABI call (pinned host AoS arrays, H2D + kernel + D2H inside
the timed region);  e2e_fused = the same matrices through the fused colvars host entry (3x3
correlation matrix up, leading eigenpair down);  roofline = algorithmic bytes / measured
launch time vs the measured HBM copy bandwidth;  strong_scaling = ONE 16 Mi batch split over
the N GPUs;  metric_sizes = n = 3 and n = 32 (the other sizes of BASELINE's metric) at this N;
other_workloads (N = 1) = the configured batches of configs[0], [2], [3] in fp64 and fp32, each
with its own clocks, the reference's measured rotation count r_ref as flop-roofline
denominator and, where it applies, cuSOLVER's syevjBatched beside it;  cpu_baseline = the
unmodified reference header (oracle/_ref) on this box's host cores over a bounded sample.
The oracle is used here only as that timed CPU baseline / rotation count.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MAT = 4                     # matrix size of the headline workload
PER_GPU_BATCH = 1 << 24       # 16M matrices per GPU
SEED = 4321
KIND_QUAT = 1
SORT_DEC = 1
MAX_SWEEPS = 50


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, scaled
    per matrix (profiles/r2_tier1_n4_traffic.json, else round 1's); None if absent."""
    for name in ("r2_tier1_n4_traffic.json", "r1_tier1_n4_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)
        except Exception:
            continue
    return None


class ClockSampler:
    """Polls NVML (SM clock, power, clock-event reasons) from a thread DURING the timed region.
    In-process polling every ~2 ms: a timed region of a few tens of ms is too short for
    `nvidia-smi -lms`, whose first sample arrives after its own start-up."""

    def __init__(self, cuda_index: int):
        import threading
        self.samples = []          # (t, sm_mhz, power_w, reasons_mask)
        self.marks = [None, None]  # perf_counter bounds of the timed region
        self.max_mhz = None
        self._stop = threading.Event()
        self.err = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception as e:  # no NVML: report that, never invent clocks
            self.err = f"nvml unavailable: {e}"
            self.thread = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = 0.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = 0
                self.samples.append((time.perf_counter(), float(mhz), pw, int(rs)))
            except Exception as e:
                self.err = str(e)
                return
            time.sleep(0.002)

    def mark(self, which):
        self.marks[which] = time.perf_counter()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "nvml unavailable"]}
        self._stop.set()
        self.thread.join(timeout=2)
        nv = self.nv
        t0, t1 = self.marks
        inside = [s for s in self.samples if t0 is not None and t1 is not None and t0 <= s[0] <= t1]
        used = inside or self.samples
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
                 "hw_power_brake": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown}
        reasons = sorted(k for k, bit in names.items() if any(s[3] & bit for s in used))
        return {"sm_mhz": statistics.median(s[1] for s in used), "sm_max_mhz": self.max_mhz,
                "samples_in_timed_region": len(inside), "samples_total": len(self.samples),
                "power_w_max": max(s[2] for s in used), "reasons": reasons}


def _bind_to_gpu_numa_node(cuda_index):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask) BEFORE the pinned host
    buffers of the e2e arm are allocated, so that they land on the GPU's NUMA node and the
    PCIe copies do not cross the socket interconnect.  Best effort: silently skipped when NVML
    cannot tell."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def time_device_steps(fn, steps, warmup, barrier, sampler=None):
    """W warm-up steps, then exactly K steps between barrier+synchronize; returns ms total."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark(0)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if sampler:
        sampler.mark(1)
    barrier()
    return e0.elapsed_time(e1)


def fp_peak(jpd, torch, elem=8):
    """Measured FMA-pipe peak (TFLOP/s) from the library's register-only FMA probe."""
    scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib = jpd.lib()
    blocks, threads, iters = 148 * 4, 256, 20000
    lib.jpd_fma_probe(elem, blocks, threads, 200, scratch.data_ptr(), st)
    torch.cuda.synchronize()
    best = 1e30
    n = 0
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        n = lib.jpd_fma_probe(elem, blocks, threads, iters, scratch.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n / (best * 1e-3) / 1e12


def host_threads():
    # all host cores this process may run on, stated explicitly: torchrun exports
    # OMP_NUM_THREADS=1, which would otherwise pin the reference arm to one thread
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(sample, passes, n=N_MAT, kind=KIND_QUAT, seed=SEED, dtype="float64"):
    """cpu_baseline leg: time the reference's own CPU implementation (oracle/_ref when it was
    built from /root/reference, else the pinned C restatement) on `sample` matrices."""
    import numpy as np
    from oracle import binding as ob
    mat = ob.fill_synthetic(n, sample, 0, kind, seed, 0, np.dtype(dtype))
    fn = ob.ref_diagonalize if ob.ref_available() else ob.oracle_diagonalize
    kind_name = "reference" if ob.ref_available() else "port"
    n_thr = host_threads()
    fn(mat[: max(1, min(sample, 100000 // (n * n)))], SORT_DEC, True, MAX_SWEEPS, n_thr)  # warm
    times, used, iters = [], 1, None
    for _ in range(passes):
        t0 = time.perf_counter()
        ev, vec, iters, used = fn(mat, SORT_DEC, True, MAX_SWEEPS, n_thr)
        times.append(time.perf_counter() - t0)
    return {"times": times, "cores": used, "kind": kind_name, "sample": sample,
            "mean_n_iters": float(iters.mean())}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    sample = args.ref_sample
    r = cpu_reference_run(sample, args.warmup + args.steps)
    timed = r["times"][args.warmup:]
    total = sum(timed)
    value = sample * len(timed) / total
    line = {
        "impl": "reference", "metric": "matrices/sec", "value": value, "unit": "matrices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),   # identical to the CUDA arm's config
        "cpu_baseline": {"value": value, "unit": "matrices/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"each step = the first {sample} matrices of the workload (bounded sample; "
                                   f"throughput is per matrix on the same generator), {len(timed)} timed passes, "
                                   f"all host threads (OpenMP), mean n_iters {r['mean_n_iters']:.2f}"},
        "e2e": {"value": value, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {
        "workload": "BASELINE configs[1]: 16M 4x4 fp64 quaternion-superposition (colvars optimal-rotation) "
                    "matrices per GPU, Diagonalize with evecs, SORT_DECREASING_EVALS, max_num_sweeps=50",
        "n": N_MAT, "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * n_gpus,
        "layout_device": "soa-interleaved", "layout_host_e2e": "aos",
        "sharding": f"{n_gpus} independent contiguous slices, no collective",
        "l2": "inputs larger than L2 (1.3 GB read + 2.7 GB written per step per GPU)",
        "generator": f"splitmix64 counter stream, seed {SEED}, kind quaternion",
    }


# ------------------------------------------------------------------ side workloads (context lines)
def reference_rotations(n, dtype, sample):
    """r_ref of SURVEY 8(d): mean rotations the reference itself performs on the first `sample`
    matrices of the same batch (= mean(n_iters) - 1): the implementation-independent work the
    flop roofline credits.  Part of the cpu_baseline leg (runs the CPU reference)."""
    import numpy as np
    from oracle import binding as ob
    A = ob.fill_synthetic(n, sample, 0, 0, 1234, 0, np.dtype(dtype))
    fn = ob.ref_diagonalize if ob.ref_available() else ob.oracle_diagonalize
    t0 = time.perf_counter()
    _, _, iters, used = fn(A, SORT_DEC, True, MAX_SWEEPS, host_threads())
    dt = time.perf_counter() - t0
    return float(iters.mean()) - 1.0, sample / dt, used


def cusolver_syevj_batched(torch, mat_aos, n, batch):
    """Informative second bar (NOT the reference): NVIDIA's own batched Jacobi,
    cusolverDn{D,S}syevjBatched, on the same device-resident matrices (default tolerance,
    100 sweeps max, eigenvectors on).  Returns matrices/s or a string saying why not."""
    import ctypes
    try:
        lib = None
        for name in ("libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"):
            try:
                lib = ctypes.CDLL(name); break
            except OSError:
                continue
        if lib is None:
            return "libcusolver not loadable"
        dbl = mat_aos.dtype == torch.float64
        pre = "cusolverDnD" if dbl else "cusolverDnS"
        h, info_p = ctypes.c_void_p(), ctypes.c_void_p()
        if lib.cusolverDnCreate(ctypes.byref(h)) != 0:
            return "cusolverDnCreate failed"
        lib.cusolverDnSetStream(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        lib.cusolverDnCreateSyevjInfo(ctypes.byref(info_p))
        a = mat_aos.clone()                                  # overwritten with the eigenvectors
        w = torch.empty((batch, n), dtype=a.dtype, device="cuda")
        dev_info = torch.empty((batch,), dtype=torch.int32, device="cuda")
        lwork = ctypes.c_int(0)
        JOBZ_VECTOR, UPLO_UPPER = 1, 1                       # CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER
        args = [h, JOBZ_VECTOR, UPLO_UPPER, n, ctypes.c_void_p(a.data_ptr()), n, ctypes.c_void_p(w.data_ptr())]
        if getattr(lib, pre + "syevjBatched_bufferSize")(*args, ctypes.byref(lwork), info_p, batch) != 0:
            return "bufferSize failed"
        work = torch.empty((max(1, lwork.value),), dtype=a.dtype, device="cuda")
        fn = getattr(lib, pre + "syevjBatched")

        def run():
            a.copy_(mat_aos)
            return fn(*args, ctypes.c_void_p(work.data_ptr()), lwork.value, ctypes.c_void_p(dev_info.data_ptr()), info_p, batch)
        if run() != 0:
            return "syevjBatched failed"
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
        c0.record(); a.copy_(mat_aos); c1.record()
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) - c0.elapsed_time(c1)       # minus the input refresh
        lib.cusolverDnDestroySyevjInfo(info_p); lib.cusolverDnDestroy(h)
        return batch / (ms * 1e-3)
    except Exception as e:  # context only: never let it break the bench
        return f"unavailable: {e}"


# n, dtype, batch (BASELINE configs), layout, timed passes, reference sample for r_ref, cuSOLVER batch (0 = skip)
SIDE_CASES = (
    (3, "float64", 1 << 23, "soa", 5, 200000, 1 << 16),      # configs[0] shape at 8 Mi (larger than L2)
    (4, "float32", 1 << 24, "soa", 5, 200000, 0),
    (16, "float64", 1 << 18, "aos", 3, 4000, 0),             # configs[2]
    (16, "float32", 1 << 18, "aos", 3, 4000, 0),
    (32, "float64", 1 << 18, "aos", 2, 1000, 1 << 12),
    (32, "float32", 1 << 18, "aos", 2, 1000, 0),
    (128, "float64", 8192, "aos", 1, 32, 0),                 # configs[3]
    (128, "float32", 8192, "aos", 1, 32, 0),                 # same shape in fp32 (context: the tier-3 block-ordering kernel)
    (256, "float64", 2048, "aos", 1, 8, 64),
    (4, "float64", 1 << 22, "soa", 3, 0, 1 << 16),            # cuSOLVER column for the headline shape
    (512, "float64", 64, "aos", 1, 16, 0),                   # beyond the configs: reference benchmarks to n = 1000
    (1000, "float64", 16, "aos", 1, 0, 0),
)


def side_workloads(jpd, torch, hbm_peak, cuda_index, quick=False):
    """The other sizes the metric and the configs name, at the configured batches, device
    resident, one GPU, each with its own clock record: context lines inside the JSON."""
    out = {}
    peaks = {"float64": fp_peak(jpd, torch, 8), "float32": fp_peak(jpd, torch, 4)}
    out["fp64_fma_peak_tflops_measured"] = round(peaks["float64"], 2)
    out["fp32_fma_peak_tflops_measured"] = round(peaks["float32"], 2)
    for n, dtype, batch, layout_name, reps, ref_sample, cus_batch in SIDE_CASES:
        if quick and n > 32:
            continue
        tdt = torch.float64 if dtype == "float64" else torch.float32
        elem = 8 if dtype == "float64" else 4
        layout = jpd.LAYOUT_SOA if layout_name == "soa" else jpd.LAYOUT_AOS
        shape = (n, n, batch) if layout == jpd.LAYOUT_SOA else (batch, n, n)
        key = f"n{n}_{'fp64' if elem == 8 else 'fp32'}"
        try:
            if n <= 256:
                m = torch.empty(shape, dtype=tdt, device="cuda")
                jpd.fill_synthetic(m, n, batch, layout, KIND_QUAT if (n == 4) else 0, SEED if n == 4 else 1234)
            else:   # beyond the counter generator's sizes: seeded U(-1,1) upper triangle, mirrored (AoS)
                g = torch.Generator(device="cuda"); g.manual_seed(1234 + n)
                m = torch.rand((batch, n, n), dtype=tdt, device="cuda", generator=g) * 2 - 1
                m = (torch.triu(m) + torch.triu(m, 1).transpose(1, 2)).contiguous()
            res = jpd.diagonalize_batched(m, layout=layout)
            torch.cuda.synchronize()
            sampler = ClockSampler(cuda_index)
            ms = time_device_steps(lambda: jpd.diagonalize_batched(m, layout=layout, out=res), reps,
                                   3 if n <= 32 else 1, lambda: None, sampler) / reps
            clocks = sampler.stop()
        except Exception as e:
            out[key] = {"error": str(e)}
            continue
        mats = batch / (ms * 1e-3)
        by = jpd.algorithmic_bytes(n, elem, True)
        entry = {"batch": batch, "layout": layout_name, "tier": int(jpd.lib().jpd_kernel_tier(n, elem, 1)),
                 "matrices_per_s": mats, "ms_per_pass": ms, "failed": int((res[2] == 0).sum().item()),
                 "mean_n_iters": float(res[2].float().mean().item()), "clocks": clocks}
        hbm_bound = hbm_peak * 1e9 / by
        if n > 256:
            # the CPU reference needs ~10 s (n = 512) / ~100 s (n = 1000) per matrix and core: timed on
            # `ref_sample` matrices of this very batch when that fits the bench's time budget
            entry["note"] = "reference published: 5.39 s (n=500) / 66.3 s (n=1000) per matrix, fp32, 1 core i5-4210U (BASELINE.md)"
            if ref_sample:
                from oracle import binding as ob
                import numpy as np
                A = m[:ref_sample].cpu().numpy()
                fn = ob.ref_diagonalize if ob.ref_available() else ob.oracle_diagonalize
                t0 = time.perf_counter()
                rev, _, rit, cores = fn(A, SORT_DEC, True, MAX_SWEEPS, host_threads())
                dt = time.perf_counter() - t0
                entry.update({"cpu_reference_matrices_per_s": ref_sample / dt, "cpu_cores": cores, "cpu_sample": ref_sample,
                              "r_ref_measured": float(rit.mean()) - 1.0,
                              "max_eval_diff_vs_cpu_reference": float((res[0][:ref_sample].cpu().numpy() - rev).__abs__().max()
                                                                      / np.abs(rev).max())})
        elif ref_sample and n != 4:
            r_ref, cpu_rate, cores = reference_rotations(n, dtype, ref_sample)
            fp_bound = peaks[dtype] * 1e12 / (r_ref * (12 * n + 4))
            bound = min(hbm_bound, fp_bound)
            entry.update({"r_ref_measured": r_ref, "r_ref_sample": ref_sample,
                          "bound": "hbm" if hbm_bound < fp_bound else ("fp64" if elem == 8 else "fp32"),
                          "roofline_matrices_per_s": bound, "frac": mats / bound,
                          "cpu_reference_matrices_per_s": cpu_rate, "cpu_cores": cores})
        elif n == 4:
            entry.update({"bound": "hbm", "roofline_matrices_per_s": hbm_bound, "frac": mats / hbm_bound})
        if cus_batch:
            aos = m[:cus_batch] if layout == jpd.LAYOUT_AOS else m[:, :, :cus_batch].permute(2, 0, 1).contiguous()
            entry["cusolver_syevjBatched_matrices_per_s"] = cusolver_syevj_batched(torch, aos.contiguous(), n, cus_batch)
            entry["cusolver_batch"] = cus_batch
        out[key] = entry
        del m, res
        torch.cuda.empty_cache()
    return out


def layout_adapter_bandwidth(jpd, torch, hbm_peak):
    """GB/s of the AoS <-> SoA adapter (read + written bytes / time) against the HBM copy peak."""
    out = {}
    for n, batch in ((3, 1 << 23), (4, 1 << 23), (32, 1 << 17)):
        a = torch.empty((batch, n, n), dtype=torch.float64, device="cuda")
        jpd.fill_synthetic(a, n, batch, jpd.LAYOUT_AOS, 0, 7)
        for name, src, sl, dl in (("aos_to_soa", a, jpd.LAYOUT_AOS, jpd.LAYOUT_SOA),):
            b = jpd.convert_layout(src, sl, dl, n, batch)
            torch.cuda.synchronize()
            ms = time_device_steps(lambda: jpd.convert_layout(src, sl, dl, n, batch), 5, 3, lambda: None) / 5
            gbs = 2.0 * a.numel() * 8 / (ms * 1e-3) / 1e9
            out[f"n{n}_{name}"] = {"GBps": gbs, "frac_of_hbm_copy_peak": gbs / hbm_peak}
            back = jpd.convert_layout(b, dl, sl, n, batch)
            torch.cuda.synchronize()
            ms = time_device_steps(lambda: jpd.convert_layout(b, dl, sl, n, batch), 5, 3, lambda: None) / 5
            gbs = 2.0 * a.numel() * 8 / (ms * 1e-3) / 1e9
            out[f"n{n}_soa_to_aos"] = {"GBps": gbs, "frac_of_hbm_copy_peak": gbs / hbm_peak,
                                       "round_trip_exact": bool(torch.equal(back, a))}
            del b, back
        del a
        torch.cuda.empty_cache()
    return out


def run_own_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import jacobi_pd_b200 as jpd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    _bind_to_gpu_numa_node(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            dist.barrier()

    def max_over_ranks(x):
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, batch = N_MAT, args.batch
    first, count = jpd.slice_of(batch * world, world, rank)
    assert count == batch
    hbm_peak, peak_src = _peaks()

    # ---- device-resident arm: SoA-interleaved batch in HBM, one launch per step
    mat = torch.empty((n, n, batch), dtype=torch.float64, device="cuda")
    jpd.fill_synthetic(mat, n, batch, jpd.LAYOUT_SOA, KIND_QUAT, SEED, b0=first)
    out = jpd.diagonalize_batched(mat, layout=jpd.LAYOUT_SOA, sort_criteria=SORT_DEC, max_num_sweeps=MAX_SWEEPS)
    torch.cuda.synchronize()
    failed = int((out[2] == 0).sum().item())

    def step():
        jpd.diagonalize_batched(mat, layout=jpd.LAYOUT_SOA, sort_criteria=SORT_DEC, calc_evecs=True,
                                max_num_sweeps=MAX_SWEEPS, out=out)

    launches0 = jpd.lib().jpd_launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total = time_device_steps(step, args.steps, args.warmup, barrier, sampler)
    launches = jpd.lib().jpd_launch_count() - launches0 - args.warmup
    clocks = sampler.stop() if sampler else None
    ms_total = max_over_ranks(ms_total)
    ms_per_step = ms_total / args.steps
    value = batch * world / (ms_per_step * 1e-3)
    launch_ms = ms_per_step  # one kernel per step, back to back on one stream
    alg_bytes = jpd.algorithmic_bytes(n, 8, True) * batch
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    traffic = _ncu_traffic()

    # ---- strong scaling of the same workload (configs[1]: "16M ... sharded over 1/2/4/8"):
    # the 16 Mi matrices of ONE global batch split into N slices of 16 Mi / N; at N = 8 a step
    # is ~170 us per GPU, so launch skew across ranks is inside the number (max over ranks)
    strong = None
    if world > 1:
        sfirst, scount = jpd.slice_of(args.batch, world, rank)
        smat = mat[:, :, :scount]   # any 16Mi/N matrices of the stream: same statistics per slice
        smat = smat.contiguous()
        jpd.fill_synthetic(smat, n, scount, jpd.LAYOUT_SOA, KIND_QUAT, SEED, b0=sfirst)
        sout = jpd.diagonalize_batched(smat, layout=jpd.LAYOUT_SOA, sort_criteria=SORT_DEC, max_num_sweeps=MAX_SWEEPS)
        torch.cuda.synchronize()
        ssteps = args.steps * 4
        sms = time_device_steps(lambda: jpd.diagonalize_batched(smat, layout=jpd.LAYOUT_SOA, sort_criteria=SORT_DEC,
                                                                max_num_sweeps=MAX_SWEEPS, out=sout),
                                ssteps, args.warmup, barrier)
        sms = max_over_ranks(sms) / ssteps
        strong = {"global_batch": args.batch, "per_gpu_batch": scount, "steps": ssteps, "ms_per_step": sms,
                  "value": args.batch / (sms * 1e-3), "unit": "matrices/s", "scaling": "strong",
                  "note": "value / (N=1 line's value) is the strong-scaling speed-up; efficiency = that / N"}
        del smat, sout
    else:
        strong = {"global_batch": args.batch, "per_gpu_batch": batch, "steps": args.steps, "ms_per_step": ms_per_step,
                  "value": value, "unit": "matrices/s", "scaling": "strong", "note": "N = 1: identical to the headline line"}

    # ---- the other two sizes of the metric (n = 3 and n = 32, fp64), weak, at every N
    metric_sizes = {}
    for mn, mb, mlayout, reps in ((3, 1 << 23, jpd.LAYOUT_SOA, 5), (32, 1 << 17, jpd.LAYOUT_AOS, 2)):
        shape = (mn, mn, mb) if mlayout == jpd.LAYOUT_SOA else (mb, mn, mn)
        mm = torch.empty(shape, dtype=torch.float64, device="cuda")
        jpd.fill_synthetic(mm, mn, mb, mlayout, 0, 1234, b0=rank * mb)
        mo = jpd.diagonalize_batched(mm, layout=mlayout)
        torch.cuda.synchronize()
        ms = max_over_ranks(time_device_steps(lambda: jpd.diagonalize_batched(mm, layout=mlayout, out=mo), reps, 3, barrier)) / reps
        metric_sizes[f"n{mn}_fp64"] = {"per_gpu_batch": mb, "matrices_per_s": mb * world / (ms * 1e-3), "ms_per_step": ms,
                                       "layout": "soa" if mlayout == jpd.LAYOUT_SOA else "aos",
                                       "failed_rank0": int((mo[2] == 0).sum().item())}
        del mm, mo
    torch.cuda.empty_cache()

    # ---- end-to-end arm: pinned host AoS buffers through the host entry point of the C ABI
    e2e = None
    e2e_fused = None
    if not args.no_e2e:
        eb = args.e2e_batch
        h_mat = torch.empty((eb, n, n), dtype=torch.float64).pin_memory()
        h_ev = torch.empty((eb, n), dtype=torch.float64).pin_memory()
        h_vec = torch.empty((eb, n, n), dtype=torch.float64).pin_memory()
        h_it = torch.empty((eb,), dtype=torch.int32).pin_memory()
        # fill host input from the same generator (device fill of an AoS buffer, copied back once)
        tmp = torch.empty((eb, n, n), dtype=torch.float64, device="cuda")
        jpd.fill_synthetic(tmp, n, eb, jpd.LAYOUT_AOS, KIND_QUAT, SEED, b0=first)
        h_mat.copy_(tmp); del tmp
        torch.cuda.synchronize()

        def timed_host(fn, steps, reps=2):
            """wall clock of `steps` calls, max over ranks; the better of `reps` back-to-back measurements
            (a host-side pipeline is exposed to scheduling noise of the box: one descheduled feeder thread
            costs 20 % of a 0.3 s measurement)"""
            for _ in range(max(1, min(args.warmup, 2))):
                fn()
            best = None
            for _ in range(reps):
                barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(steps):
                    fn()
                torch.cuda.synchronize()
                dt_ = max_over_ranks(time.perf_counter() - t0)
                best = dt_ if best is None else min(best, dt_)
            return best

        def e2e_step():
            jpd.diagonalize_batched_host(h_mat, h_ev, h_vec, h_it, n, eb, jpd.LAYOUT_AOS, SORT_DEC, True,
                                         MAX_SWEEPS, device=local_rank)

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        dt = timed_host(e2e_step, e2e_steps)

        # context: the same byte volumes as plain concurrent pinned copies, ALL ranks at once
        # (the PCIe / host-memory bound of this arm on this box at this N)
        def copy_bound(up_elems, down_elems):
            d_in = torch.empty((up_elems,), dtype=torch.float64, device="cuda")
            d_out = torch.empty((down_elems,), dtype=torch.float64, device="cuda")
            hh_in = torch.empty((up_elems,), dtype=torch.float64).pin_memory()
            hh_out = torch.empty((down_elems,), dtype=torch.float64).pin_memory()
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            best = 1e30
            for _ in range(3):
                barrier(); torch.cuda.synchronize(); t1 = time.perf_counter()
                with torch.cuda.stream(s_up):
                    d_in.copy_(hh_in, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    hh_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
                best = min(best, max_over_ranks(time.perf_counter() - t1))
            return eb * world / best

        pcie_bound = None if args.no_extras else copy_bound(eb * n * n, eb * (n + n * n) + (eb + 1) // 2)
        e2e = {"value": eb * world * e2e_steps / dt, "unit": "matrices/s",
               "h2d_bytes_per_step": int(h_mat.numel() * 8),
               "d2h_bytes_per_step": int(h_ev.numel() * 8 + h_vec.numel() * 8 + h_it.numel() * 4),
               "steps": e2e_steps, "timing": "wall clock of `steps` host calls, better of 2 measurements, max over ranks",
               "batch_per_gpu": eb, "host_layout": "aos", "host_memory": "pinned",
               "checksum_evals": float(h_ev[:: max(1, eb // 4096)].sum().item()),
               "pcie_copy_bound": pcie_bound,
               "frac_of_pcie_copy_bound": (eb * world * e2e_steps / dt / pcie_bound) if pcie_bound else None}
        lam_check = h_ev[:4096, 0].clone()
        del h_mat, h_ev, h_vec

        # ---- the same matrices through the FUSED host entry (what colvars actually needs):
        # 3x3 correlation matrix up (72 B), leading eigenvalue + quaternion + n_iters down (44 B)
        h_corr = torch.empty((eb, 9), dtype=torch.float64).pin_memory()
        h_lam = torch.empty((eb,), dtype=torch.float64).pin_memory()
        h_quat = torch.empty((eb, 4), dtype=torch.float64).pin_memory()
        tmp = torch.empty((eb, 9), dtype=torch.float64, device="cuda")
        jpd.fill_synthetic(tmp, 3, eb, jpd.LAYOUT_AOS, 2, SEED, b0=first)   # kind 2: the correlation matrices of kind 1
        h_corr.copy_(tmp); del tmp
        torch.cuda.synchronize()

        def fused_step():
            jpd.quaternion_from_correlation_host(h_corr, h_lam, h_quat, h_it, eb, jpd.LAYOUT_AOS, False, MAX_SWEEPS,
                                                 device=local_rank)

        dtf = timed_host(fused_step, e2e_steps)
        fused_bound = None if args.no_extras else copy_bound(eb * 9, eb * 5 + (eb + 1) // 2)
        e2e_fused = {"value": eb * world * e2e_steps / dtf, "unit": "matrices/s",
                     "entry": "jpd_quaternion_from_correlation_host_f64 (leading eigenpair)",
                     "h2d_bytes_per_step": int(h_corr.numel() * 8),
                     "d2h_bytes_per_step": int(h_lam.numel() * 8 + h_quat.numel() * 8 + h_it.numel() * 4),
                     "steps": e2e_steps, "batch_per_gpu": eb, "host_memory": "pinned",
                     "same_leading_eigenvalues_as_e2e": bool(torch.equal(h_lam[:4096], lam_check)),
                     "pcie_copy_bound": fused_bound,
                     "frac_of_pcie_copy_bound": (eb * world * e2e_steps / dtf / fused_bound) if fused_bound else None}
        del h_corr, h_lam, h_quat, h_it

        # ---- the metric's other sizes end to end (pinned AoS host buffers through the same host entry)
        for mn, mb in ((3, 1 << 23), (32, 1 << 17)):
            hm = torch.empty((mb, mn, mn), dtype=torch.float64).pin_memory()
            hev = torch.empty((mb, mn), dtype=torch.float64).pin_memory()
            hvec = torch.empty((mb, mn, mn), dtype=torch.float64).pin_memory()
            hit = torch.empty((mb,), dtype=torch.int32).pin_memory()
            tmp = torch.empty((mb, mn, mn), dtype=torch.float64, device="cuda")
            jpd.fill_synthetic(tmp, mn, mb, jpd.LAYOUT_AOS, 0, 1234, b0=rank * mb)
            hm.copy_(tmp); del tmp
            torch.cuda.synchronize()
            msteps = max(1, min(e2e_steps, 3))
            dts = timed_host(lambda: jpd.diagonalize_batched_host(hm, hev, hvec, hit, mn, mb, jpd.LAYOUT_AOS, SORT_DEC, True,
                                                                  MAX_SWEEPS, device=local_rank), msteps)
            metric_sizes[f"n{mn}_fp64"]["e2e"] = {
                "value": mb * world * msteps / dts, "unit": "matrices/s", "steps": msteps,
                "h2d_bytes_per_step": int(hm.numel() * 8),
                "d2h_bytes_per_step": int(hev.numel() * 8 + hvec.numel() * 8 + hit.numel() * 4),
                "failed_rank0": int((hit == 0).sum().item())}
            del hm, hev, hvec, hit

    line = None
    if rank == 0:
        line = {
            "metric": "matrices/sec", "value": value, "unit": "matrices/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_matrix": jpd.algorithmic_bytes(n, 8, True),
                         "kernel": "jpd_small_kernel<4,double,SOA,EVECS>",
                         "traffic": (traffic["dram_bytes_per_matrix"] * batch) if traffic else None,
                         "traffic_source": traffic["source"] if traffic else None},
            "e2e": e2e, "e2e_fused": e2e_fused, "gpu_launches": int(launches), "clocks": clocks,
            "not_converged": failed, "strong_scaling": strong, "metric_sizes": metric_sizes,
        }
    if rank == 0 and world == 1 and not args.no_extras:
        del mat, out
        torch.cuda.empty_cache()
        line["other_workloads"] = side_workloads(jpd, torch, hbm_peak, local_rank, quick=args.quick)
        line["layout_adapter"] = layout_adapter_bandwidth(jpd, torch, hbm_peak)
        r = cpu_reference_run(args.ref_sample, 3)
        best = min(r["times"])
        line["cpu_baseline"] = {"value": r["sample"] / best, "unit": "matrices/s", "cores": r["cores"],
                                "kind": r["kind"],
                                "sample": f"first {r['sample']} matrices of the workload, best of 3 passes, "
                                          f"OpenMP over all host threads, mean n_iters {r['mean_n_iters']:.2f}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["own", "reference"], default="own")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="matrices per GPU (default 16M)")
    ap.add_argument("--e2e-batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-sample", type=int, default=1 << 22, help="matrices per CPU-reference pass")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--quick", action="store_true", help="side workloads only up to n = 32")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
