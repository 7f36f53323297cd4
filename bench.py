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
the timed region);  roofline = algorithmic bytes / measured launch time vs the measured HBM
copy bandwidth;  cpu_baseline = the unmodified reference header (oracle/_ref) on this box's
host cores over a bounded sample.  The oracle is used here only as that timed baseline.
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
    per matrix (profiles/r1_tier1_n4_traffic.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_tier1_n4_traffic.json")) as f:
            return json.load(f)
    except Exception:
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


def side_workloads(jpd, torch, hbm_peak):
    """The other sizes the metric and the configs name (n=3 thread tier, n=16/32 warp tier,
    n=128 CTA tier, n=256 cluster tier; fp64), device-resident, one GPU: context lines inside
    the JSON, not the headline."""
    out = {}
    fp64_tflops = fp_peak(jpd, torch, 8)
    out["fp64_fma_peak_tflops_measured"] = round(fp64_tflops, 2)
    # mean rotations of the reference on these inputs (r_ref) -> flop roofline
    r_ref = {3: 8.71, 16: 486.0, 32: 2078.0, 128: 35388.8, 256: 143937.8}
    cases = ((3, 1 << 23, jpd.LAYOUT_SOA, 5), (16, 1 << 18, jpd.LAYOUT_AOS, 5),
             (32, 1 << 17, jpd.LAYOUT_AOS, 5), (128, 2048, jpd.LAYOUT_AOS, 2), (256, 528, jpd.LAYOUT_AOS, 1))
    for n, batch, layout, reps in cases:
        shape = (n, n, batch) if layout == jpd.LAYOUT_SOA else (batch, n, n)
        m = torch.empty(shape, dtype=torch.float64, device="cuda")
        jpd.fill_synthetic(m, n, batch, layout, 0, 1234)
        res = jpd.diagonalize_batched(m, layout=layout)
        torch.cuda.synchronize()
        ms = time_device_steps(lambda: jpd.diagonalize_batched(m, layout=layout, out=res), reps, 3, lambda: None) / reps
        mats = batch / (ms * 1e-3)
        by = jpd.algorithmic_bytes(n, 8, True)
        hbm_bound = hbm_peak * 1e9 / by
        fp_bound = fp64_tflops * 1e12 / (r_ref[n] * (12 * n + 4))
        bound = min(hbm_bound, fp_bound)
        out[f"n{n}_fp64"] = {
            "batch": batch, "layout": "soa" if layout else "aos", "tier": int(jpd.lib().jpd_kernel_tier(n, 8, 1)),
            "matrices_per_s": mats, "ms_per_pass": ms, "bound": "hbm" if hbm_bound < fp_bound else "fp64",
            "roofline_matrices_per_s": bound, "frac": mats / bound,
            "failed": int((res[2] == 0).sum().item()),
        }
        del m, res
    return out


def cpu_reference_run(sample, passes):
    """Time the reference's own CPU implementation on `sample` matrices of the workload."""
    import numpy as np
    from oracle import binding as ob
    mat = ob.fill_synthetic(N_MAT, sample, 0, KIND_QUAT, SEED, 0, np.float64)
    fn = ob.ref_diagonalize if ob.ref_available() else ob.oracle_diagonalize
    kind = "reference" if ob.ref_available() else "port"
    # all host cores this process may run on, stated explicitly: torchrun exports
    # OMP_NUM_THREADS=1, which would otherwise pin the reference arm to one thread
    try:
        n_thr = len(os.sched_getaffinity(0))
    except AttributeError:
        n_thr = os.cpu_count() or 1
    fn(mat[: min(sample, 100000)], SORT_DEC, True, MAX_SWEEPS, n_thr)  # warm
    times, used, iters = [], 1, None
    for _ in range(passes):
        t0 = time.perf_counter()
        ev, vec, iters, used = fn(mat, SORT_DEC, True, MAX_SWEEPS, n_thr)
        times.append(time.perf_counter() - t0)
    return {"times": times, "cores": used, "kind": kind, "sample": sample,
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
        "config": workload_config(args.gpus, sample_note=f"each step = {sample} matrices of the workload (bounded sample)"),
        "cpu_baseline": {"value": value, "unit": "matrices/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": f"{sample} matrices x {len(timed)} passes, all host threads (OpenMP), "
                                   f"mean n_iters {r['mean_n_iters']:.2f}"},
        "e2e": {"value": value, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, sample_note=None):
    cfg = {
        "workload": "BASELINE configs[1]: 16M 4x4 fp64 quaternion-superposition (colvars optimal-rotation) "
                    "matrices per GPU, Diagonalize with evecs, SORT_DECREASING_EVALS, max_num_sweeps=50",
        "n": N_MAT, "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * n_gpus,
        "layout_device": "soa-interleaved", "layout_host_e2e": "aos",
        "sharding": f"{n_gpus} independent contiguous slices, no collective",
        "l2": "inputs larger than L2 (1.3 GB read + 2.7 GB written per step per GPU)",
        "generator": f"splitmix64 counter stream, seed {SEED}, kind quaternion",
    }
    if sample_note:
        cfg["sample"] = sample_note
    return cfg


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

    # ---- end-to-end arm: pinned host AoS buffers through the host entry point of the C ABI
    e2e = None
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

        def e2e_step():
            jpd.diagonalize_batched_host(h_mat, h_ev, h_vec, h_it, n, eb, jpd.LAYOUT_AOS, SORT_DEC, True,
                                         MAX_SWEEPS, device=local_rank)

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        # context: the same byte volumes as plain concurrent pinned copies (the PCIe bound of this arm)
        pcie_bound = None
        if rank == 0 and world == 1 and not args.no_extras:
            d_in = torch.empty((eb, n, n), dtype=torch.float64, device="cuda")
            d_out = torch.empty((eb * (n + n * n) + (eb + 1) // 2,), dtype=torch.float64, device="cuda")
            h_out = torch.empty(d_out.shape, dtype=torch.float64).pin_memory()
            s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
            best = 1e30
            for _ in range(3):
                torch.cuda.synchronize(); t1 = time.perf_counter()
                with torch.cuda.stream(s_up):
                    d_in.copy_(h_mat, non_blocking=True)
                with torch.cuda.stream(s_dn):
                    h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t1)
            pcie_bound = eb / best
            del d_in, d_out, h_out
        e2e = {"value": eb * world * e2e_steps / dt, "unit": "matrices/s",
               "h2d_bytes_per_step": int(h_mat.numel() * 8),
               "d2h_bytes_per_step": int(h_ev.numel() * 8 + h_vec.numel() * 8 + h_it.numel() * 4),
               "steps": e2e_steps, "batch_per_gpu": eb, "host_layout": "aos", "host_memory": "pinned",
               "checksum_evals": float(h_ev[:: max(1, eb // 4096)].sum().item()),
               "pcie_copy_bound": pcie_bound,
               "frac_of_pcie_copy_bound": (eb * e2e_steps / dt / pcie_bound) if pcie_bound else None}
        del h_mat, h_ev, h_vec, h_it

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
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "not_converged": failed,
        }
    if rank == 0 and world == 1 and not args.no_extras:
        del mat, out
        torch.cuda.empty_cache()
        line["other_workloads"] = side_workloads(jpd, torch, hbm_peak)
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
