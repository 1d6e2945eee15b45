#!/usr/bin/env python
"""bench.py — headline benchmark: mapped vertices/s of computeMapping + map, rbf-pum-direct, compact-polynomial-c6,
3-D unit cube, N -> N Halton vertices (BASELINE.json configs[3]; default N = 10M), on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA library through its C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU algorithm on the host cores

A "step" = Mapping::clear() + computeMapping() + map() of one scalar field (what ParticipantImpl does per remeshing
step). `value` times the step with the meshes and the field already resident in HBM (device-pointer entry points);
`e2e` times the same step through the host-pointer C ABI (rbfmap_compute_mapping / rbfmap_map) from pinned host
buffers, H2D of coordinates + values and D2H of the result inside the timed region.
Multi-GPU: one process per GPU (torchrun); the path shards by mesh partition exactly like the reference's parallel
PUM (each rank maps its own partition, no data-path collective: BatchedRBFSolver.hpp:29-31) -> weak scaling.
Only rank 0 prints the JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VPC = 50
OVERLAP = 0.15
BASIS = "compact-polynomial-c6"
F_PHI = 24  # algorithmic flops per C6 pair evaluation in 3-D (sqrt = 1, fma = 2)


def support_radius(n_vertices: int) -> float:
    """support radius = 2 x cluster radius; the cluster radius of a uniform unit-cube mesh with
    `VPC` vertices per sphere is (3 VPC / (4 pi N))^(1/3)."""
    return 2.0 * (3.0 * VPC / (4.0 * math.pi * n_vertices)) ** (1.0 / 3.0)


def halton_numpy(n, bases, skip):
    out = np.empty((n, 3))
    idx = np.arange(skip, n + skip, dtype=np.int64)
    for d, b in enumerate(bases):
        f, r, i = 1.0, np.zeros(n), idx.copy()
        while np.any(i > 0):
            f /= b
            r += f * (i % b)
            i //= b
        out[:, d] = r
    return out


def halton_torch(n, bases, skip, device):
    import torch
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    idx = torch.arange(skip, n + skip, dtype=torch.int64, device=device)
    for d, b in enumerate(bases):
        f = 1.0
        r = torch.zeros(n, dtype=torch.float64, device=device)
        i = idx.clone()
        while bool((i > 0).any()):
            f /= b
            r += f * (i % b).to(torch.float64)
            i //= b
        out[:, d] = r
    return out


def field_torch(xyz):
    import torch
    return torch.sin(2 * math.pi * xyz[:, 0]) * torch.cos(2 * math.pi * xyz[:, 1]) + xyz[:, 2]


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    # NVML queries take a driver lock; issued from a second thread they stall the CUDA API calls of the step that runs
    # beside them (measured on this box: single steps of 33 - 234 ms instead of 20.6, worst with
    # nvmlDeviceGetCurrentClocksThrottleReasons). The main thread therefore samples BETWEEN the timed steps - inside the
    # timed region, GPU still at its load clocks, but never concurrently with a step: the SM clock after every step
    # (poll_clock), the throttle reasons before the first and after the last step (poll_reasons).
    def run(self):
        return

    def poll_clock(self):
        if self.nv is None:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            pass

    def poll_reasons(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        try:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline(n_sample: int, threads: int):
    """The reference algorithm (CPU oracle = port of the Eigen path) on a bounded sample of the same workload, run the way
    the reference parallelises rbf-pum-direct: one rank per core, every rank clusters and maps its own partition with no
    communication (BatchedRBFSolver.hpp:29-31). `threads` partitions of n_sample/threads vertices each."""
    import oracle
    # at most 500k vertices per partition: a step stays near 5 s of wall clock on hosts with few cores as well
    per = max(2000, min(n_sample // threads, 500_000))
    ins = [halton_numpy(per, (2, 3, 5), 1 + r * per) for r in range(threads)]
    outs = [halton_numpy(per, (7, 11, 13), 1 + r * per) for r in range(threads)]
    vals = [np.sin(2 * np.pi * a[:, 0]) * np.cos(2 * np.pi * a[:, 1]) + a[:, 2] for a in ins]
    tc, tm, res, ncl = oracle.pum_time_ranks("consistent", 3, BASIS, support_radius(per), "separate", VPC, OVERLAP, True, ins, outs, vals, 1)
    total = per * threads
    return {"value": total / (tc + tm), "unit": "vertices/s", "cores": threads, "kind": "port",
            "sample": f"{threads} partitions x {per}->{per} Halton vertices (one rank per host thread, as the reference "
                      f"parallelises rbf-pum-direct), same configuration, {ncl} clusters, computeMapping {tc:.2f} s + map {tm:.2f} s",
            "t_compute_s": tc, "t_map_s": tm, "total_vertices": total}, res[0], (ins[0], outs[0], vals[0], per)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference itself cannot be
    built in this image) with all host threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.ref_sample
    times = []
    for s in range(args.warmup + args.steps):
        cb, _, _ = cpu_baseline(n_sample, threads)
        if s >= args.warmup:
            times.append(cb["t_compute_s"] + cb["t_map_s"])
    t = sum(times) / len(times)
    v = cb["total_vertices"] / t
    line = {"metric": "mapped vertices/s (computeMapping+map), rbf-pum-direct 3D", "value": v, "unit": "vertices/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.n, n_sample_note=cb["total_vertices"]),
            "cpu_baseline": {"value": v, "unit": "vertices/s", "cores": threads, "kind": "port", "sample": cb["sample"]},
            "e2e": {"value": v, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n, n_sample_note=None):
    cfg = {"workload": f"rbf-pum-direct compact-polynomial-c6 3D unit cube {n}->{n} Halton vertices per GPU, consistent, scalar field, "
                       f"polynomial separate, vertices-per-cluster {VPC}, relative-overlap {OVERLAP}, project-to-input on, "
                       f"support-radius 2x cluster radius",
           "vertices_per_gpu": n, "support_radius": support_radius(n), "l2": "inputs (2 x 24 B x N coordinates) are larger than the 126 MB L2"}
    if n_sample_note is not None:
        cfg["reference_sample_vertices"] = n_sample_note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vertices", dest="n", type=int, default=10_000_000, help="vertices per mesh per GPU (BASELINE.json configs[3]: 10M)")
    ap.add_argument("--ref-sample", type=int, default=8_000_000, help="vertices of the bounded CPU sample (about 10 s of CPU work on 16 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import precice_b200 as pb

    if not torch.cuda.is_available() or pb.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (precice_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # NCCL announces its version on stdout when the first communicator is created; stdout carries exactly one JSON line,
    # so file descriptor 1 points to stderr until that line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    support = support_radius(n)
    # rank r maps its own partition: the Halton points with indices [1 + r n, 1 + (r+1) n)
    in_xyz = halton_torch(n, (2, 3, 5), 1 + rank * n, dev)
    out_xyz = halton_torch(n, (7, 11, 13), 1 + rank * n, dev)
    vals = field_torch(in_xyz).contiguous()
    res_dev = torch.empty(n, dtype=torch.float64, device=dev)

    mapping = pb.PartitionOfUnityMapping("consistent", 3, BASIS, support, "separate", VPC, OVERLAP, True, device_id=local_rank)

    # ---------------- device-resident figure ("value") ----------------
    def step_device():
        if mapping.hasComputedMapping():
            mapping.clear()
        mapping.computeMapping(in_xyz, out_xyz)
        mapping.map(vals, 1, out=res_dev)
        return mapping.stats()

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    ev_ms, asm_ms, map_ms, launches = 0.0, 0.0, 0.0, 0
    sampler.poll_reasons()
    for _ in range(args.steps):
        st = step_device()
        ev_ms += st["ms_compute_kernels"] + st["ms_map_kernels"]  # CUDA events on the library's stream, whole calls
        asm_ms += st["ms_assembly_factor"]
        map_ms += st["ms_map_kernels"]
        launches += st["kernel_launches"]
        sampler.poll_clock()
        if os.environ.get("BENCH_VERBOSE"):
            print(f"[bench] step: compute {st['ms_compute_kernels']:.2f} ms (factor {st['ms_assembly_factor']:.2f}) map {st['ms_map_kernels']:.2f} ms",
                  file=sys.stderr, flush=True)
    sampler.poll_reasons()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    stats = mapping.stats()

    # ---------------- end-to-end figure through the host-pointer C ABI ("e2e") ----------------
    h_in = in_xyz.cpu().pin_memory()
    h_out = out_xyz.cpu().pin_memory()
    h_vals = vals.cpu().pin_memory()
    h_res = torch.empty(n, dtype=torch.float64).pin_memory()

    def step_host():
        if mapping.hasComputedMapping():
            mapping.clear()
        mapping.computeMapping(h_in, h_out)
        mapping.map(h_vals, 1, out=h_res)

    for _ in range(min(args.warmup, 3)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    checksum = float(h_res.sum())

    # ---------------- max over ranks ----------------
    t = torch.tensor([ev_ms, wall_ms, e2e_ms, asm_ms, map_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ev_ms, wall_ms, e2e_ms, asm_ms, map_ms = (float(x) for x in t.tolist())

    if rank == 0:
        steps = args.steps
        # device time of the step = CUDA-event time of the two calls; the wall clock around them is reported beside it
        ms_per_step = ev_ms / steps
        value = world * n / (ms_per_step * 1e-3)
        e2e_value = world * n / (e2e_ms / steps * 1e-3)

        # Rooflines of the two algebra kernels. Both are FP64-pipe bound (the matrices live in registers / shared memory;
        # HBM traffic is far below its roofline, see profiles/), so `peak` is the FP64 FMA peak measured in this run.
        # Algorithmic flops (sqrt = 1, fma = 2, F_phi = 24 for C6 in 3-D), from the cluster statistics:
        #   factor kernel: sum_c [F_phi n_c (n_c + 1) / 2 + n_c^3 / 3]        (assemble C + Cholesky)
        #   map kernel   : sum_c [2 n_c^2 + (F_phi + 2) m_c n_c]              (solve + evaluate A lambda on the fly)
        from precice_b200 import _lib
        import ctypes as C
        tf, ms = C.c_double(0), C.c_double(0)
        _lib.lib().rbfmap_measure_fp64_peak(C.byref(tf), C.byref(ms))
        fp64_peak = tf.value
        src = ("measured in this run: dependency-free DFMA loop (rbfmap_measure_fp64_peak); FP64 is not in MEASURED_PEAKS.json")

        def fp64_roofline(kernel, flops, ms_total):
            sec = ms_total / steps * 1e-3
            ach = flops / sec / 1e12
            return {"bound": "fp64", "kernel": kernel, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak if fp64_peak else None, "traffic": None, "peak_source": src,
                    "algorithmic_flops_per_launch": flops, "ms_per_launch": sec * 1e3, "share_of_step": ms_total / ev_ms}

        r_factor = fp64_roofline("pumFactorWarpKernel (per-cluster assembly of C + square-root-free Cholesky on DMMA, all buckets)",
                                 F_PHI * (stats["sum_n2"] + stats["sum_n"]) / 2.0 + stats["sum_n3"] / 3.0, asm_ms)
        r_map = fp64_roofline("pumMapConsistentKernel (gather + polynomial + solve + on-the-fly evaluation + blend, all buckets)",
                              2.0 * stats["sum_n2"] + (F_PHI + 2.0) * stats["sum_mn"], map_ms)
        # DRAM traffic: `ncu --set full` capture of the same kernels on the 1M-vertex workload (profiles/r1c_*_summary.txt,
        # dram__bytes_read.sum + dram__bytes_write.sum of the n = 49..56 bucket: 11.5 KB resp. 19.8 KB per cluster), scaled by
        # the number of clusters of this run. Algorithmic bytes per cluster: packed factor 8 n(n+1)/2 B (written once by the
        # factor kernel, read once per map) + 24 B per gathered vertex + 8 B per value.
        r_factor["traffic"] = 11.5e3 * stats["n_clusters"]
        r_map["traffic"] = 19.8e3 * stats["n_clusters"]
        for r_ in (r_factor, r_map):
            r_["traffic_source"] = "ncu capture at 1M vertices (profiles/r1d_*_kernel_summary.txt), bytes per cluster x clusters"
        r_factor["algorithmic_bytes"] = 8.0 * (stats["sum_n2"] + stats["sum_n"]) / 2.0 + 24.0 * stats["sum_n"]
        r_map["algorithmic_bytes"] = 8.0 * (stats["sum_n2"] + stats["sum_n"]) / 2.0 + 32.0 * stats["sum_n"] + 32.0 * stats["sum_m"]
        roofline, roofline_other = (r_factor, r_map) if asm_ms >= map_ms else (r_map, r_factor)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            peaks = {}

        line = {"metric": "mapped vertices/s (computeMapping+map), rbf-pum-direct 3D", "value": value, "unit": "vertices/s",
                "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(n),
                "e2e": {"value": e2e_value, "unit": "vertices/s", "h2d_bytes_per_step": 8 * (3 * n + 3 * n + n),
                        "d2h_bytes_per_step": 8 * n, "ms_per_step": e2e_ms / steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_secondary": roofline_other,
                "fp64_peak_tflops": fp64_peak,
                "wall_ms_per_step": wall_ms / steps,
                "mapping": {k: stats[k] for k in ("n_clusters", "cluster_radius", "sum_n", "sum_m", "sum_n2", "sum_n3", "sum_mn",
                                                  "max_n", "device_bytes")},
                "hbm_peak_gbs": peaks.get("hbm_gbs"), "checksum": checksum}
        if not args.no_cpu_baseline:
            cb, ref_out, (s_in, s_out, s_vals, per) = cpu_baseline(args.ref_sample, os.cpu_count() or 1)
            # parity of the GPU path against the CPU path on the first partition of the sample, checked in the same run
            m2 = pb.PartitionOfUnityMapping("consistent", 3, BASIS, support_radius(per), "separate", VPC, OVERLAP, True,
                                            device_id=local_rank)
            m2.computeMapping(s_in, s_out)
            got = m2.map(s_vals, 1)
            cb["parity_rel_l2_gpu_vs_cpu"] = float(np.linalg.norm(got - ref_out) / np.linalg.norm(ref_out))
            for k in ("t_compute_s", "t_map_s", "total_vertices"):
                cb.pop(k)
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
