#!/usr/bin/env python
"""bench.py — headline benchmark: mapped vertices/s of computeMapping + map, rbf-pum-direct, compact-polynomial-c6,
3-D unit cube, N -> N Halton vertices (BASELINE.json configs[3]; default N = 10M IN TOTAL), on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA library through its C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU algorithm on the host cores

A "step" = Mapping::clear() + computeMapping() + map() of one scalar field (what ParticipantImpl does per remeshing
step). `value` times the step with the meshes and the field already resident in HBM (device-pointer entry points);
`e2e` times the same step through the host-pointer C ABI (rbfmap_compute_mapping / rbfmap_map) from pinned host
buffers, H2D of coordinates + values and D2H of the result inside the timed region. On N > 1 GPUs every rank sends 1/N
of the meshes and values over PCIe (all-gather over NVLink) and reads back 1/N of the complete result
(rbfmap_map_sliced: reduce-scatter over NVLink); each rank's slice is checked against the device-resident result.

Multi-GPU (torchrun, one process per GPU): ONE 10M -> 10M mapping is sharded over the ranks -> "scaling": "strong".
Every rank holds the replicated meshes, repeats the cheap clustering of the whole mesh (`replicated_ms`, the Amdahl
term) and processes a slab of the clusters (include/rbfmap.h): `--shard-mode duplicate` (default; boundary
clusters on both sides of a cut, no collective in map) or `exchange` (one owner per cluster, NCCL all-reduce of weight
sums and partial blends); the other mode and the round-1 weak-scaling figure (N independent 10M cubes) are reported
under `secondary`. Only rank 0 prints the JSON line.

The result of the timed mapping is checked in every run against tests/golden/pum_cfg4_10m.npz: the CPU oracle on the
same 10M meshes at 10 000 sampled vertices (consistent + conservative, scalar + 3-vector), see `parity`.
Secondary legs (`secondary`): BASELINE.json configs[1] (50k rbf-global-direct), configs[2] (500k rbf-global-iterative,
sharded over the ranks), configs[4] (2M-vertex surface remeshing), each with its own roofline fraction.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VPC = 50
OVERLAP = 0.15
BASIS = "compact-polynomial-c6"
F_PHI = 24  # algorithmic flops per C6 pair evaluation in 3-D (sqrt = 1, fma = 2)
GOLDEN_10M = os.path.join(ROOT, "tests", "golden", "pum_cfg4_10m.npz")
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "traffic_10m.json")
PARITY_TOL = 1e-10  # north_star: within 1e-10 relative L2 of the reference (direct)


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


def vector_field_torch(xyz):
    """(f, 2f + 1, -f), vertex-major interleaved."""
    import torch
    f = field_torch(xyz)
    return torch.stack([f, 2 * f + 1, -f], dim=1).contiguous().reshape(-1)


class ClockSampler:
    """SM clock and throttle reasons of one GPU during the timed region (pynvml).

    NVML queries take a driver lock; issued from a second thread they stalled the CUDA API calls of the step that ran
    beside them (measured: single steps of 33 - 234 ms instead of 20.6, worst with
    nvmlDeviceGetCurrentClocksThrottleReasons). The main thread therefore samples BETWEEN the timed steps - inside the
    timed region, GPU still at its load clocks, but never concurrently with a step: the SM clock after every step
    (poll_clock), the throttle reasons before the first and after the last step (poll_reasons)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

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

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline(n_sample: int, threads: int):
    """The reference algorithm (CPU oracle = port of the Eigen path) on a bounded sample of the same workload, run the way
    the reference parallelises rbf-pum-direct: one rank per core, every rank clusters and maps its own partition with no
    communication (BatchedRBFSolver.hpp:29-31). `threads` partitions of n_sample/threads vertices each. PUM is O(N), so
    the per-vertex rate of the sample is what the reference would reach on the 10M mesh split over the same ranks."""
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
                      f"parallelises rbf-pum-direct; {total} vertices in total, extrapolated linearly in N to the 10M workload: "
                      f"partition of unity is O(N)), same configuration, {ncl} clusters, computeMapping {tc:.2f} s + map {tm:.2f} s",
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
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.n, args.gpus, None, n_sample_note=cb["total_vertices"]),
            "cpu_baseline": {"value": v, "unit": "vertices/s", "cores": threads, "kind": "port", "sample": cb["sample"]},
            "e2e": {"value": v, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n, world, shard_mode, n_sample_note=None):
    cfg = {"workload": f"rbf-pum-direct compact-polynomial-c6 3D unit cube {n}->{n} Halton vertices IN TOTAL (one mapping"
                       + (f" sharded over {world} GPUs, shard mode {shard_mode}" if world > 1 and shard_mode else "")
                       + f"), consistent, scalar field, polynomial separate, vertices-per-cluster {VPC}, relative-overlap {OVERLAP}, "
                         f"project-to-input on, support-radius 2x cluster radius",
           "total_vertices": n, "support_radius": support_radius(n),
           "l2": "inputs (2 x 24 B x N coordinates) are larger than the 126 MB L2"}
    if n_sample_note is not None:
        cfg["reference_sample_vertices"] = n_sample_note
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off (what `numactl --cpunodebind` does in a production
    launcher), BEFORE any host buffer is allocated: pinned staging memory is then first-touched on the socket whose PCIe
    root the GPU sits on, and eight ranks do not all stream through one socket's memory controllers. Returns a note for
    the JSON line. BENCH_NUMA=0 switches it off; silently skipped where sysfs has no topology (containers)."""
    if os.environ.get("BENCH_NUMA", "1") == "0":
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        dev = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"
        node = int(open(dev + "/numa_node").read())
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return "no topology"
        os.sched_setaffinity(0, cpus)
        return f"node {node} ({len(cpus)} cpus)"
    except Exception as e:  # no NVML / sysfs: leave the affinity alone
        return f"unavailable ({type(e).__name__})"


class Ctx:
    """Per-process state: device, process group, helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.affinity0 = os.sched_getaffinity(0)  # restored for the CPU baseline, which uses every host thread
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, tensor):
        if self.world > 1:
            self.dist.all_reduce(tensor)
        return tensor

    def join(self, mapping, sharded=True):
        """rbfmap_distributed_init: rank 0 creates the NCCL id, torch.distributed broadcasts it."""
        import precice_b200 as pb
        if self.world > 1 and sharded:
            ident = [pb.nccl_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(ident, src=0)
            mapping.distributedInit(ident[0], self.rank, self.world)
        return mapping

    def pum(self, constraint, n, shard_mode, sharded=True, execution_mode="minimal-memory"):
        import precice_b200 as pb
        m = pb.PartitionOfUnityMapping(constraint, 3, BASIS, support_radius(n), "separate", VPC, OVERLAP, True, device_id=self.local_rank,
                                       shard_mode=shard_mode, execution_mode=execution_mode)
        return self.join(m, sharded)


def complete_result(ctx, mapping, res):
    """The complete mapped field on every rank: duplicate-mode shards fill only the vertices they own (the others are
    zero), so the field is the sum over the ranks; every other configuration already returns it."""
    st = mapping.stats()
    if ctx.world > 1 and st["shard_axis"] >= 0 and mapping._cfg.shard_mode == 0 and mapping._cfg.constraint != 1:
        res = res.clone()
        ctx.sum_over_ranks(res)
    return res


def golden_parity(ctx, n, shard_mode, res_scalar, mapping, in_xyz, out_xyz):
    """The timed mapping (and three more untimed ones) against the oracle fixture on the same 10M meshes."""
    torch = ctx.torch
    if n != 10_000_000 or not os.path.exists(GOLDEN_10M):
        return {"fixture": None, "note": "the committed fixture covers the default 10M workload only"}
    g = np.load(GOLDEN_10M)
    ids = torch.from_numpy(g["samples"].astype(np.int64)).to(ctx.dev)

    def rel(res, key, dims):
        got = res.reshape(n, dims)[ids].cpu().numpy()
        ref = g[key]
        return float(np.linalg.norm(got - ref) / np.linalg.norm(ref))

    st = mapping.stats()
    out = {"fixture": "tests/golden/pum_cfg4_10m.npz (CPU oracle on the same 10M meshes, 10 000 sampled vertices; "
                      "tests/golden/make_golden_10m.py)", "tolerance_rel_l2": PARITY_TOL,
           "n_clusters": int(st["n_clusters_global"]), "n_clusters_oracle": int(g["consistent_n_clusters"]),
           "cluster_radius_equal": bool(st["cluster_radius"] == float(g["consistent_radius"])),
           "consistent_scalar": rel(complete_result(ctx, mapping, res_scalar), "consistent_1", 1)}
    v3 = vector_field_torch(in_xyz)
    timings = {}  # library event times of these (second, warm) calls: the other directions / data dimensions of the workload

    def timed_map(m, values, dims, key):
        m.map(values, dims)
        res = m.map(values, dims)
        (timings[key],) = ctx.max_over_ranks([m.stats()["ms_map_kernels"]])
        return res

    out["consistent_vector3"] = rel(complete_result(ctx, mapping, timed_map(mapping, v3, 3, "consistent_vector3_ms_map")), "consistent_3", 3)
    mc = ctx.pum("conservative", n, shard_mode)
    mc.computeMapping(in_xyz, out_xyz)
    mc.clear()
    mc.computeMapping(in_xyz, out_xyz)
    stc = mc.stats()
    (timings["conservative_ms_compute_mapping"],) = ctx.max_over_ranks([stc["ms_compute_kernels"]])
    out["conservative_n_clusters"] = int(stc["n_clusters_global"])
    out["conservative_n_clusters_oracle"] = int(g["conservative_n_clusters"])
    out["conservative_scalar"] = rel(timed_map(mc, field_torch(in_xyz).contiguous(), 1, "conservative_scalar_ms_map"), "conservative_1", 1)
    out["conservative_vector3"] = rel(timed_map(mc, v3, 3, "conservative_vector3_ms_map"), "conservative_3", 3)
    out["timings_ms"] = timings
    del mc
    out["ok"] = bool(all(out[k] <= PARITY_TOL for k in ("consistent_scalar", "consistent_vector3", "conservative_scalar",
                                                        "conservative_vector3"))
                     and out["n_clusters"] == out["n_clusters_oracle"] and out["cluster_radius_equal"]
                     and out["conservative_n_clusters"] == out["conservative_n_clusters_oracle"])
    return out


def time_steps(ctx, mapping, compute_args, vals, res, steps, warmup, sampler=None):
    """K timed steps (clear + computeMapping + map); device time = CUDA events of the library around its two calls."""
    def step():
        if mapping.hasComputedMapping():
            mapping.clear()
        mapping.computeMapping(*compute_args)
        mapping.map(vals, 1, out=res)
        return mapping.stats()

    for _ in range(warmup):
        step()
    ctx.barrier()
    t0 = time.perf_counter()
    acc = {"ev": 0.0, "asm": 0.0, "map": 0.0, "rep": 0.0, "coll": 0.0, "launches": 0}
    if sampler:
        sampler.poll_reasons()
        ctx.barrier()  # the NVML query above delays the ranks by different amounts
    for _ in range(steps):
        st = step()
        acc["ev"] += st["ms_compute_kernels"] + st["ms_map_kernels"]
        acc["asm"] += st["ms_assembly_factor"]
        acc["map"] += st["ms_map_kernels"]
        acc["rep"] += st["ms_replicated"]
        acc["coll"] += st["ms_collectives"]
        acc["launches"] += st["kernel_launches"]
        if sampler:
            # NVML queries take 1 - 3 ms and delay the rank that issues them; in a sharded step the other ranks would
            # then wait for it inside the next computeMapping (collectives) and that wait would be booked as device time of
            # the step. The ranks therefore line up again after the query; the wall-clock figure contains all of it.
            sampler.poll_clock()
            if ctx.world > 1:
                ctx.dist.barrier()
        if os.environ.get("BENCH_VERBOSE"):
            print(f"[bench] rank {ctx.rank} step: compute {st['ms_compute_kernels']:.2f} ms (replicated {st['ms_replicated']:.2f}, factor "
                  f"{st['ms_assembly_factor']:.2f}) map {st['ms_map_kernels']:.2f} ms", file=sys.stderr, flush=True)
    if sampler:
        sampler.poll_reasons()
    ctx.barrier()
    acc["wall"] = (time.perf_counter() - t0) * 1e3
    return acc, mapping.stats()


def time_host_steps(ctx, mapping, h_in, h_out, h_vals, h_res, steps, warmup, sliced=False):
    """`sliced`: every rank reads back its 1/world slice of the complete result (rbfmap_map_sliced), the mirror of the
    upload in which every rank sends 1/world of the inputs; returns (ms, slice begin, slice end)."""
    span = [0, h_res.numel()]

    def step():
        t0 = time.perf_counter()
        if mapping.hasComputedMapping():
            mapping.clear()
        mapping.computeMapping(h_in, h_out)
        t1 = time.perf_counter()
        if sliced:
            _, span[0], span[1] = mapping.mapSliced(h_vals, 1, out=h_res)
        else:
            mapping.map(h_vals, 1, out=h_res)
        if os.environ.get("BENCH_VERBOSE"):
            print(f"[bench] rank {ctx.rank} host step: clear + computeMapping {1e3 * (t1 - t0):.2f} ms, map {1e3 * (time.perf_counter() - t1):.2f} ms",
                  file=sys.stderr, flush=True)

    for _ in range(warmup):
        step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ctx.barrier()
    return (time.perf_counter() - t0) * 1e3, span[0], span[1]


# ---------------------------------------------------------------------------------------------------------------------
def leg_cfg2(ctx, fp64_peak):
    """BASELINE.json configs[1]: rbf-global-direct, compact-polynomial-c2, 3-D unit cube, 50k -> 50k, consistent scalar +
    conservative 3-vector, one GPU (the global-direct solve stays on one device: rank 0 only)."""
    import precice_b200 as pb
    torch = ctx.torch
    n, support = 50_000, 0.2
    inp = halton_torch(n, (2, 3, 5), 1, ctx.dev)
    out = halton_torch(n, (7, 11, 13), 1, ctx.dev)
    f = field_torch(inp).contiguous()
    res = {"workload": f"rbf-global-direct compact-polynomial-c2 support-radius {support}, 3D unit cube {n}->{n} Halton vertices, polynomial "
                       f"separate, 1 GPU (replicas only: the dense solve does not shard)"}
    m = pb.RadialBasisFctMapping("consistent", 3, "compact-polynomial-c2", support, polynomial="separate", device_id=ctx.local_rank)
    t_c, t_f = [], []
    for _ in range(2):  # first pass = warm-up
        if m.hasComputedMapping():
            m.clear()
        m.computeMapping(inp, out)
        st = m.stats()
        t_c.append(st["ms_compute_kernels"])
        t_f.append(st["ms_assembly_factor"])
    m.map(f, 1)
    r = m.map(f, 1)
    ms_map = m.stats()["ms_map_kernels"]
    exact = field_torch(out)
    # size-independent parity property at the full BASELINE size (the dense CPU oracle cannot factorise 50k x 50k in a bench
    # run): with a separated polynomial a linear field is reproduced exactly (RadialBasisFctSolver.hpp:441-468)
    lin = lambda x: 1.0 + 2.0 * x[:, 0] - x[:, 1] + 3.0 * x[:, 2]
    rl = m.map(lin(inp).contiguous(), 1)
    res["linear_reproduction_rel_l2"] = float(torch.linalg.norm(rl - lin(out)) / torch.linalg.norm(lin(out)))
    res.update({"ms_compute_mapping": t_c[-1], "ms_assembly_factorisation": t_f[-1], "ms_map_consistent_scalar": ms_map,
                "vertices_per_s": n / ((t_c[-1] + ms_map) * 1e-3),
                "interpolation_error_rel_l2": float(torch.linalg.norm(r - exact) / torch.linalg.norm(exact))})
    flops = n ** 3 / 3.0 + 16.0 * n * (n + 1) / 2.0  # Cholesky + assembly of the lower triangle (F_phi = 16 for C2)
    ach = flops / (t_f[-1] * 1e-3) / 1e12
    res["roofline"] = {"bound": "fp64 (DMMA trailing updates)", "kernel": "assembleKernel + potrfInvKernel + dgemmNtKernel (Cholesky of C)",
                       "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                       "algorithmic_flops": flops}
    del m
    mc = pb.RadialBasisFctMapping("conservative", 3, "compact-polynomial-c2", support, polynomial="separate", device_id=ctx.local_rank)
    mc.computeMapping(inp, out)
    v3 = vector_field_torch(inp)
    mc.map(v3, 3)
    r3 = mc.map(v3, 3)
    res["ms_map_conservative_vector3"] = mc.stats()["ms_map_kernels"]
    s_in, s_out = v3.reshape(n, 3).sum(0), r3.reshape(n, 3).sum(0)
    res["conservation_rel_error"] = float(((s_in - s_out).abs() / s_in.abs().clamp(min=1.0)).max())
    del mc
    return res


def leg_cfg3(ctx, fp64_peak):
    """BASELINE.json configs[2]: rbf-global-iterative, gaussian with support 0.05, 3-D unit cube, 500k -> 500k, matrix-free
    CG sharded over the ranks (NCCL all-gather of the direction vector + all-reduce of the dot products per iteration)."""
    import precice_b200 as pb
    torch = ctx.torch
    n, support, rtol = 500_000, 0.05, 1e-9
    shape = math.sqrt(-math.log(1e-9)) / support
    inp = halton_torch(n, (2, 3, 5), 1, ctx.dev)
    out = halton_torch(n, (7, 11, 13), 1, ctx.dev)
    f = field_torch(inp).contiguous()

    def solve(sharded):
        m = pb.RadialBasisFctMapping("consistent", 3, "gaussian", shape, polynomial="separate", solver="iterative", solver_rtol=rtol,
                                     device_id=ctx.local_rank)
        ctx.join(m, sharded)
        best = None
        for _ in range(2):  # the second pass is timed; clear() resets the warm start so that it solves from zero again
            if m.hasComputedMapping():
                m.clear()
            m.computeMapping(inp, out)
            if sharded:
                ctx.barrier()
            r = m.map(f, 1)
            st = m.stats()
            best = (st["ms_map_kernels"], st["ms_compute_kernels"], st["iterations"], st["residual"], st)
        del m
        return r, best

    r_sh, (ms_map, ms_comp, its, resid, st_sh) = solve(True)
    ms_map, ms_comp = ctx.max_over_ranks([ms_map, ms_comp])
    exact = field_torch(out)
    interp_err = float(torch.linalg.norm(r_sh - exact) / torch.linalg.norm(exact))
    res = {"workload": f"rbf-global-iterative gaussian shape {shape:.2f} (support {support}), 3D unit cube {n}->{n} Halton vertices, "
                       f"solver-rtol {rtol}, CG on the assembled sparse system, rows sharded over {ctx.world} GPU(s)",
           "n_gpus": ctx.world, "iterations": its, "relative_residual": resid, "ms_compute_mapping": ms_comp, "ms_map": ms_map,
           "ms_per_iteration": ms_map / max(its, 1), "vertices_per_s": n / ((ms_comp + ms_map) * 1e-3),
           "interpolation_error_rel_l2": interp_err,
           "note": "computeMapping includes the assembly of the sparse system and the CG solve for the rescaling coefficients "
                   "(PetRadialBasisFctMapping.hpp:470-490)"}
    if ctx.world > 1:
        r_1, (ms1, _, its1, _, _) = solve(False)
        res["rel_l2_sharded_vs_single_gpu"] = float(torch.linalg.norm(r_sh - r_1) / torch.linalg.norm(r_1))
        res["ms_map_single_gpu"] = ms1
        res["iterations_single_gpu"] = its1
    # Every CG iteration streams this rank's rows of the assembled system matrix once: 12 B per non-zero (8 B value + 4 B
    # column) plus the vectors; HBM-bound. (Round 1 re-evaluated the basis function in every iteration: 2.3 ms / iteration.)
    nnz = st_sh["sum_mn"]
    bytes_it = 12.0 * nnz + 8.0 * 6 * n / ctx.world
    gbs = bytes_it / (ms_map / max(its, 1) * 1e-3) / 1e9
    peaks = _peaks()
    res["nonzeros_rank0"] = nnz
    res["roofline"] = {"bound": "hbm", "kernel": "spmvKernel (assembled system matrix of this rank's rows, one product per CG iteration)",
                       "achieved": gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                       "algorithmic_bytes_per_iteration": bytes_it,
                       "note": "whole iteration (product + 3 vector kernels + collectives) of rank 0 against the HBM copy bandwidth"}
    return res


def _xorshift_unit(ctx, n, seed):
    """xorshift32 stream per vertex ('xorshift seed 42'): n x 3 numbers in [-1, 1)."""
    torch = ctx.torch
    m32 = 0xFFFFFFFF
    x = (torch.arange(1, 3 * n + 1, dtype=torch.int64, device=ctx.dev) * 2654435761 + seed * 40503) & m32
    x = torch.where(x == 0, torch.full_like(x, 0x9E3779B9), x)
    for _ in range(2):
        x = x ^ ((x << 13) & m32)
        x = x ^ (x >> 17)
        x = x ^ ((x << 5) & m32)
    return (x.to(torch.float64) / 2147483648.0 - 1.0).reshape(n, 3)


def fibonacci_sphere(ctx, n, rotate):
    torch = ctx.torch
    i = torch.arange(n, dtype=torch.float64, device=ctx.dev)
    z = 1.0 - (2.0 * i + 1.0) / n
    r = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    phi = i * (math.pi * (3.0 - math.sqrt(5.0)))
    p = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=1)
    if rotate:  # fixed rotation about the axis (1, 2, 3) / sqrt(14) by 0.7 rad (Rodrigues)
        k = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device=ctx.dev) / math.sqrt(14.0)
        a = 0.7
        p = p * math.cos(a) + torch.cross(k.expand_as(p), p, dim=1) * math.sin(a) + k * (p @ k).unsqueeze(1) * (1.0 - math.cos(a))
    return p.contiguous()


def leg_cfg5(ctx, fp64_peak, shard_mode, steps=10, warmup=3, n=2_000_000):
    """With project-to-input the reference ALGORITHM itself can leave output vertices of a surface mesh without cluster
    (the oracle stops with the same 'could not be assigned to any cluster' error on such steps); the leg
    then repeats without the projection and says so."""
    out = _leg_cfg5(ctx, fp64_peak, shard_mode, steps, warmup, n, True)
    if "failed" in out and "UNASSIGNED" in out["failed"]:
        first = out["failed"]
        out = _leg_cfg5(ctx, fp64_peak, shard_mode, steps, warmup, n, False)
        out["project_to_input"] = False
        out["with_project_to_input"] = first
    return out


def _leg_cfg5(ctx, fp64_peak, shard_mode, steps, warmup, n, project):
    """BASELINE.json configs[4] as SURVEY §8(d) specifies it: 2M-vertex unit sphere (Fibonacci lattice), out-mesh = the
    rotated lattice, every step all vertices are perturbed by <= 0.1 h (xorshift, seed 42 + step) and clear() +
    computeMapping() + map() run again; one mapping sharded over the ranks."""
    import precice_b200 as pb
    torch = ctx.torch
    h = math.sqrt(4.0 * math.pi / n)  # mesh width of the lattice
    base_in, base_out = fibonacci_sphere(ctx, n, False), fibonacci_sphere(ctx, n, True)
    support = 2.0 * math.sqrt(VPC * 4.0 * math.pi / (math.pi * n))  # 2 x radius of a disc with VPC vertices on the unit sphere
    m = pb.PartitionOfUnityMapping("consistent", 3, BASIS, support, "separate", VPC, OVERLAP, project, device_id=ctx.local_rank, shard_mode=shard_mode)
    ctx.join(m)
    res_dev = torch.empty(n, dtype=torch.float64, device=ctx.dev)
    ev, failed = [], None
    for s in range(warmup + steps):
        xin = (base_in + 0.1 * h / math.sqrt(3.0) * _xorshift_unit(ctx, n, 42 + 2 * s)).contiguous()
        xout = (base_out + 0.1 * h / math.sqrt(3.0) * _xorshift_unit(ctx, n, 43 + 2 * s)).contiguous()
        vals = (torch.sin(3.0 * xin[:, 0]) * torch.cos(2.0 * xin[:, 1]) + torch.exp(xin[:, 2])).contiguous()
        torch.cuda.synchronize()
        if m.hasComputedMapping():
            m.clear()
        try:
            m.computeMapping(xin, xout)
        except pb.MappingError as e:
            failed = f"step {s}: {e.status}: {str(e)[:120]}"
            break
        m.map(vals, 1, out=res_dev)
        st = m.stats()
        if s >= warmup:
            ev.append(st["ms_compute_kernels"] + st["ms_map_kernels"])
    out = {"workload": f"rbf-pum-direct compact-polynomial-c6, remeshing: {n}->{n} vertices on the unit sphere (Fibonacci lattice, out-mesh "
                       f"rotated), perturbed by <= 0.1 h every step (xorshift32, seed 42 + step), clear + computeMapping + map (scalar) per "
                       f"step, one mapping sharded over {ctx.world} GPU(s)", "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
           "project_to_input": project}
    if failed:
        out["failed"] = failed
        return out
    ev.sort()
    med = ev[len(ev) // 2]
    med, mx = ctx.max_over_ranks([med, ev[-1]])
    full = complete_result(ctx, m, res_dev)
    exact = torch.sin(3.0 * xout[:, 0]) * torch.cos(2.0 * xout[:, 1]) + torch.exp(xout[:, 2])
    flops = (F_PHI * (st["sum_n2"] + st["sum_n"]) / 2.0 + st["sum_n3"] / 3.0) + (2.0 * st["sum_n2"] + (F_PHI + 2.0) * st["sum_mn"])
    ach = flops / (med * 1e-3) / 1e12
    out.update({"statistic": "median over the timed steps, max over ranks", "ms_per_step": med, "ms_per_step_max": mx,
                "vertices_per_s": n / (med * 1e-3), "ms_replicated_last_step": st["ms_replicated"],
                "interpolation_error_rel_l2": float(torch.linalg.norm(full - exact) / torch.linalg.norm(exact)),
                "mapping_rank0": {k: st[k] for k in ("n_clusters", "n_clusters_global", "cluster_radius", "sum_n", "sum_m", "max_n")},
                "roofline": {"bound": "fp64", "kernel": "whole step (clustering + membership + factorisation + map), algebra flops of rank 0",
                             "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None}})
    del m
    return out


def leg_weak_replicas(ctx, n, steps=2, warmup=2):
    """Round-1 figure kept for comparison: N independent n -> n cubes, one per rank, no data-path collective."""
    torch = ctx.torch
    in_xyz = halton_torch(n, (2, 3, 5), 1 + ctx.rank * n, ctx.dev)
    out_xyz = halton_torch(n, (7, 11, 13), 1 + ctx.rank * n, ctx.dev)
    vals = field_torch(in_xyz).contiguous()
    res = torch.empty(n, dtype=torch.float64, device=ctx.dev)
    m = ctx.pum("consistent", n, "duplicate", sharded=False)
    acc, _ = time_steps(ctx, m, (in_xyz, out_xyz), vals, res, steps, warmup)
    (ev,) = ctx.max_over_ranks([acc["ev"]])
    del m
    return {"workload": f"{ctx.world} independent {n}->{n} mappings, one per GPU (weak scaling of replicas, the round-1 N > 1 figure)",
            "ms_per_step": ev / steps, "vertices_per_s": ctx.world * n / (ev / steps * 1e-3), "scaling": "weak"}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vertices", dest="n", type=int, default=10_000_000, help="vertices per mesh IN TOTAL (BASELINE.json configs[3]: 10M)")
    ap.add_argument("--shard-mode", default="duplicate", choices=["duplicate", "exchange"])
    ap.add_argument("--ref-sample", type=int, default=8_000_000, help="vertices of the bounded CPU sample (about 10 s of CPU work on 16 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="all", help="comma list of: headline (always), e2e, parity, modes, secondary, cpu; or all")
    args = ap.parse_args()
    legs = {"e2e", "parity", "modes", "secondary", "cpu"} if args.legs == "all" else set(args.legs.split(","))
    if args.no_cpu_baseline:
        legs.discard("cpu")

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return

    import torch
    import precice_b200 as pb

    if not torch.cuda.is_available() or pb.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (precice_b200 has no CPU fallback)")
    # NCCL announces its version on stdout when the first communicator is created; stdout carries exactly one JSON line,
    # so file descriptor 1 points to stderr until that line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    ctx = Ctx(args)
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    if os.environ.get("BENCH_TRACE_RANK") is not None:  # development: phase trace (adds synchronisations) of one rank only
        os.environ["RBFMAP_TRACE"] = "1" if rank == int(os.environ["BENCH_TRACE_RANK"]) else "0"

    n = args.n
    # every rank holds the complete (replicated) meshes of the ONE mapping
    in_xyz = halton_torch(n, (2, 3, 5), 1, dev)
    out_xyz = halton_torch(n, (7, 11, 13), 1, dev)
    vals = field_torch(in_xyz).contiguous()
    res_dev = torch.empty(n, dtype=torch.float64, device=dev)
    mapping = ctx.pum("consistent", n, args.shard_mode)

    # ---------------- device-resident figure ("value") ----------------
    sampler = ClockSampler(ctx.local_rank)
    acc, stats = time_steps(ctx, mapping, (in_xyz, out_xyz), vals, res_dev, args.steps, args.warmup, sampler)
    clocks = sampler.result()
    res_timed = res_dev.clone()

    # ---------------- end-to-end figure through the host-pointer C ABI ("e2e") ----------------
    e2e_ms = e2e_pageable_ms = None
    checksum = None
    if "e2e" in legs:
        h_in, h_out, h_vals = in_xyz.cpu().pin_memory(), out_xyz.cpu().pin_memory(), vals.cpu().pin_memory()
        h_res = torch.zeros(n, dtype=torch.float64).pin_memory()
        e2e_ms, lo, hi = time_host_steps(ctx, mapping, h_in, h_out, h_vals, h_res, args.steps, min(args.warmup, 3), sliced=world > 1)
        # the slices of the ranks tile the complete field: each is checked against the device-resident result
        want = complete_result(ctx, mapping, res_timed)[lo:hi]
        got = h_res[lo:hi].to(dev)
        if hi > lo and float((got - want).norm() / want.norm().clamp_min(1e-300)) > 1e-12:
            raise SystemExit(f"bench.py: rank {rank}: the end-to-end result differs from the device-resident one")
        checksum = float(ctx.sum_over_ranks(got.sum().reshape(1)).item()) if world > 1 else float(h_res.sum())
        if world == 1:  # pageable host memory (what a caller without pinned staging buffers hands over): driver-staged copies
            p_in, p_out, p_vals, p_res = h_in.clone(), h_out.clone(), h_vals.clone(), torch.empty(n, dtype=torch.float64)
            e2e_pageable_ms = time_host_steps(ctx, mapping, p_in, p_out, p_vals, p_res, max(2, args.steps // 2), 1)[0] / max(2, args.steps // 2)
            del p_in, p_out, p_vals, p_res

    # ---------------- max over ranks ----------------
    keys = ["ev", "wall", "asm", "map", "rep", "coll"]
    red = dict(zip(keys, ctx.max_over_ranks([acc[k] for k in keys])))
    (e2e_max,) = ctx.max_over_ranks([e2e_ms or 0.0])
    steps = args.steps
    ms_per_step = red["ev"] / steps

    # ---------------- parity against the committed oracle fixture (untimed) ----------------
    parity = None
    if "parity" in legs:
        if not mapping.hasComputedMapping():
            mapping.computeMapping(in_xyz, out_xyz)
        parity = golden_parity(ctx, n, args.shard_mode, res_timed, mapping, in_xyz, out_xyz)

    # ---------------- map-only figures, both execution modes; the other shard mode ----------------
    secondary = {}
    if "modes" in legs:
        map_only = {}
        for mode in ("minimal-memory", "minimal-compute"):
            try:
                mm = ctx.pum("consistent", n, args.shard_mode, execution_mode=mode)
                mm.computeMapping(in_xyz, out_xyz)
                stc = mm.stats()
                t = []
                for _ in range(4):
                    mm.map(vals, 1, out=res_dev)
                    t.append(mm.stats()["ms_map_kernels"])
                (ms_map,) = ctx.max_over_ranks([min(t[1:])])
                (ms_comp,) = ctx.max_over_ranks([stc["ms_compute_kernels"]])
                entry = {"ms_map": ms_map, "vertices_per_s_map_only": n / (ms_map * 1e-3), "ms_compute_mapping": ms_comp,
                         "evaluation_bytes": stc["evaluation_bytes"], "device_bytes": stc["device_bytes"]}
                if mode == "minimal-compute" and stc["evaluation_bytes"]:
                    # the stored operators are streamed once per map: HBM-bound
                    peaks = _peaks()
                    gbs = stc["evaluation_bytes"] / (ms_map * 1e-3) / 1e9
                    entry["roofline"] = {"bound": "hbm", "kernel": "pumMapStoredKernel (one GEMV with the stored operator per cluster)",
                                         "achieved": gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                         "frac": gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                                         "algorithmic_bytes": stc["evaluation_bytes"]}
                    full = complete_result(ctx, mm, res_dev)
                    entry["rel_l2_vs_minimal_memory"] = float(torch.linalg.norm(full - complete_result(ctx, mapping, res_timed))
                                                              / torch.linalg.norm(full))
                map_only[mode] = entry
                del mm
            except pb.MappingError as e:
                map_only[mode] = {"error": f"{e.status}: {e}"}
        secondary["map_only"] = map_only
        # several samples in one call (rbfmap_map_batch: the stamples / data of DataContext::mapData as components of one
        # field): three scalar samples as ONE 3-component map against three separate maps, device-resident
        if not mapping.hasComputedMapping():
            mapping.computeMapping(in_xyz, out_xyz)
        v3 = vector_field_torch(in_xyz)
        t1, t3 = [], []
        for _ in range(3):
            mapping.map(vals, 1, out=res_dev)
            t1.append(mapping.stats()["ms_map_kernels"])
            mapping.map(v3, 3)
            t3.append(mapping.stats()["ms_map_kernels"])
        (ms1, ms3) = ctx.max_over_ranks([min(t1), min(t3)])
        secondary["batched_samples"] = {"ms_three_separate_scalar_maps": 3 * ms1, "ms_one_batched_map_of_three": ms3,
                                        "note": "rbfmap_map_batch interleaves the samples on the device and maps them as one field"}
        if world > 1:
            other = "exchange" if args.shard_mode == "duplicate" else "duplicate"
            mo = ctx.pum("consistent", n, other)
            acc_o, st_o = time_steps(ctx, mo, (in_xyz, out_xyz), vals, res_dev, max(2, steps // 2), 2)
            ev_o, coll_o = ctx.max_over_ranks([acc_o["ev"], acc_o["coll"]])
            k = max(2, steps // 2)
            secondary["other_shard_mode"] = {"shard_mode": other, "ms_per_step": ev_o / k, "vertices_per_s": n / (ev_o / k * 1e-3),
                                             "ms_collectives_per_map": coll_o / k, "clusters_rank0": st_o["n_clusters"]}
            del mo
            secondary["weak_replicas"] = leg_weak_replicas(ctx, n)

    # measured FP64 peak (roofline denominator of the FP64-bound kernels)
    from precice_b200 import _lib
    import ctypes as C
    tf, ms = C.c_double(0), C.c_double(0)
    _lib.lib().rbfmap_measure_fp64_peak(C.byref(tf), C.byref(ms))
    fp64_peak = tf.value

    if "secondary" in legs:
        del mapping
        mapping = None
        torch.cuda.empty_cache()
        for name, fn in (("cfg3_global_iterative_500k", lambda: leg_cfg3(ctx, fp64_peak)),
                         ("cfg5_remeshing_sphere_2m", lambda: leg_cfg5(ctx, fp64_peak, args.shard_mode))):
            try:
                secondary[name] = fn()
            except Exception as e:  # a failing secondary leg must not take the headline line with it
                secondary[name] = {"error": repr(e)[:300]}
                if world > 1:
                    raise
        if rank == 0:
            try:
                secondary["cfg2_global_direct_50k"] = leg_cfg2(ctx, fp64_peak)
            except Exception as e:
                secondary["cfg2_global_direct_50k"] = {"error": repr(e)[:300]}
        ctx.barrier()

    if rank == 0:
        value = n / (ms_per_step * 1e-3)
        src = "measured in this run: dependency-free DFMA loop (rbfmap_measure_fp64_peak); FP64 is not in MEASURED_PEAKS.json"

        # Rooflines of the two algebra kernels (rank 0's shard when sharded). Both are FP64-pipe bound (the matrices live in
        # registers / shared memory; HBM traffic is far below its roofline), so `peak` is the FP64 FMA peak measured in this run.
        # Algorithmic flops (sqrt = 1, fma = 2, F_phi = 24 for C6 in 3-D), from the cluster statistics:
        #   factor kernel: sum_c [F_phi n_c (n_c + 1) / 2 + n_c^3 / 3]        (assemble C + Cholesky)
        #   map kernel   : sum_c [2 n_c^2 + (F_phi + 2) m_c n_c]              (solve + evaluate A lambda on the fly)
        def fp64_roofline(kernel, flops, ms_total):
            sec = ms_total / steps * 1e-3
            ach = flops / sec / 1e12
            return {"bound": "fp64", "kernel": kernel, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach / fp64_peak if fp64_peak else None, "traffic": None, "peak_source": src,
                    "algorithmic_flops_per_launch": flops, "ms_per_launch": sec * 1e3, "share_of_step": ms_total / red["ev"]}

        r_factor = fp64_roofline("pumFactorWarpKernel (per-cluster assembly of C + square-root-free Cholesky on DMMA, all buckets)",
                                 F_PHI * (stats["sum_n2"] + stats["sum_n"]) / 2.0 + stats["sum_n3"] / 3.0, acc["asm"])
        r_map = fp64_roofline("pumMapConsistentKernel (gather + polynomial + solve + on-the-fly evaluation + blend, all buckets)",
                              2.0 * stats["sum_n2"] + (F_PHI + 2.0) * stats["sum_mn"], acc["map"])
        # DRAM traffic: dram__bytes_read.sum + dram__bytes_write.sum of an `ncu --set full` capture of this build on the 10M
        # workload (profiles/traffic_10m.json, written by scripts/ncu_traffic.py), per cluster x clusters of this run.
        # Algorithmic bytes per cluster: packed factor 8 n(n+1)/2 B (written once by the factor kernel, read once per
        # map) + 24 B per gathered vertex + 8 B per value.
        try:
            tr = json.load(open(TRAFFIC_JSON))
            for r_, key in ((r_factor, "factor"), (r_map, "map")):
                r_["traffic"] = tr[key]["dram_bytes_per_cluster"] * stats["n_clusters"]
                r_["traffic_source"] = tr["source"]
        except Exception:
            pass
        r_factor["algorithmic_bytes"] = 8.0 * (stats["sum_n2"] + stats["sum_n"]) / 2.0 + 24.0 * stats["sum_n"]
        r_map["algorithmic_bytes"] = 8.0 * (stats["sum_n2"] + stats["sum_n"]) / 2.0 + 32.0 * stats["sum_n"] + 32.0 * stats["sum_m"]
        roofline, roofline_other = (r_factor, r_map) if acc["asm"] >= acc["map"] else (r_map, r_factor)
        peaks = _peaks()

        line = {"metric": "mapped vertices/s (computeMapping+map), rbf-pum-direct 3D", "value": value, "unit": "vertices/s",
                "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(n, world, args.shard_mode),
                "gpu_launches": acc["launches"], "clocks": clocks, "roofline": roofline, "roofline_secondary": roofline_other,
                "fp64_peak_tflops": fp64_peak, "wall_ms_per_step": red["wall"] / steps,
                "replicated_ms": red["rep"] / steps, "collectives_ms_per_map": red["coll"] / steps,
                "ms_assembly_factorisation": red["asm"] / steps, "ms_map": red["map"] / steps,
                "mapping": {k: stats[k] for k in ("n_clusters", "n_clusters_global", "cluster_radius", "sum_n", "sum_m", "sum_n2", "sum_n3",
                                                  "sum_mn", "max_n", "device_bytes", "shard_axis", "shard_lo", "shard_hi")},
                "hbm_peak_gbs": peaks.get("hbm_gbs")}
        if e2e_ms is not None:
            line["e2e"] = {"value": n / (e2e_max / steps * 1e-3), "unit": "vertices/s",
                           "h2d_bytes_per_step": 8 * (3 * n + 3 * n + n), "d2h_bytes_per_step": 8 * n,
                           "ms_per_step": e2e_max / steps, "host_memory": "pinned (rbfmap_host_alloc / cudaHostAlloc)", "numa": ctx.numa,
                           "note": "sharded: every rank uploads 1/world of the meshes and values over PCIe (all-gather over NVLink) and "
                                   "reads back 1/world of the complete result (rbfmap_map_sliced: reduce-scatter over NVLink); bytes are "
                                   "totals over the ranks" if world > 1 else "one GPU: complete upload and read-back"}
            line["checksum"] = checksum
            if e2e_pageable_ms is not None:
                line["e2e_pageable"] = {"value": n / (e2e_pageable_ms * 1e-3), "unit": "vertices/s", "ms_per_step": e2e_pageable_ms,
                                        "host_memory": "pageable (driver-staged cudaMemcpyAsync)"}
        if parity is not None:
            line["parity"] = parity
        if secondary:
            line["secondary"] = secondary
        if "cpu" in legs:
            os.sched_setaffinity(0, ctx.affinity0)  # the CPU arm gets all host cores, not only the GPU's NUMA node
            cb, ref_out, (s_in, s_out, s_vals, per) = cpu_baseline(args.ref_sample, os.cpu_count() or 1)
            # parity of the GPU path against the CPU path on the first partition of the sample, checked in the same run
            m2 = pb.PartitionOfUnityMapping("consistent", 3, BASIS, support_radius(per), "separate", VPC, OVERLAP, True,
                                            device_id=ctx.local_rank)
            m2.computeMapping(s_in, s_out)
            got = m2.map(s_vals, 1)
            cb["parity_rel_l2_gpu_vs_cpu"] = float(np.linalg.norm(got - ref_out) / np.linalg.norm(ref_out))
            for k in ("t_compute_s", "t_map_s", "total_vertices"):
                cb.pop(k)
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        if parity is not None and parity.get("fixture") and not parity["ok"]:
            print("bench.py: PARITY FAILURE against the oracle fixture: " + json.dumps(parity), file=sys.stderr, flush=True)
            sys.exit(3)
    if world > 1:
        ctx.barrier()
        ctx.dist.destroy_process_group()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


if __name__ == "__main__":
    main()
