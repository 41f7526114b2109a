#!/usr/bin/env python
"""bench.py — headline benchmark of the exact inner-product top-k engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nq 1250] [--rows 12000000]

A "step" = one pass of the hot path over one batch of synthetic queries: exact IP top-100 of
`nq` (default 1,250) 768-d queries against the 12M x 768 bf16 synthetic passage corpus
(BASELINE.json configs[1]).  With N GPUs the corpus is row-sharded over the ranks
(viquae_b200.sharded), every step includes the all-gather of the per-rank candidates and the
k-way merge.  Rank 0 prints ONE JSON line (see README / DESIGN.md for the keys).

`--impl reference` times the reference's CPU path instead: FAISS is not installable offline, so
it is the FAISS-equivalent CPU port in oracle/flat_ip.py (blocked sgemm on all host cores +
k-selection) on a bounded sample of the same synthetic inputs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK_ROWS = 1_000_000
CORPUS_SEED = 20240117  # chunk c uses seed CORPUS_SEED + c: identical corpus for every GPU count
QUERY_SEED = 7


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nq", type=int, default=1250)
    ap.add_argument("--rows", type=int, default=12_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16", "f32", "f32x3"])
    ap.add_argument("--variant", default=None, choices=[None, "auto", "tc_cg1", "tc_cg2", "stream"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-calls", type=int, default=0, help="also time this many single calls end to end (p50/p99, BASELINE config 5)")
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="corpus rows of the CPU sample (default: 1.5M inline, 3M for --impl reference)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- synthetic inputs
def gen_chunk(c: int, rows: int, dim: int, device, dtype):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(CORPUS_SEED + c)
    return torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32).to(dtype)


def gen_queries(nq: int, dim: int, device, dtype):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(QUERY_SEED)
    return torch.randn((nq, dim), generator=g, device=device, dtype=torch.float32).to(dtype)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.1] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU port (reference arm / cpu_baseline)
def cpu_port_sample(args, device, sample_rows: int, repeats: int = 1, warmup: int = 0):
    """Times the FAISS-equivalent CPU port on all queries x the first `sample_rows` corpus rows
    (fp32 upcast of the SAME synthetic bf16 values) and scales the time linearly to the full corpus."""
    import numpy as np
    import torch
    from oracle import flat_ip

    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32, "f32x3": torch.float32}[args.dtype]
    sample_rows = min(sample_rows, args.rows)
    parts, c, left = [], 0, sample_rows
    while left > 0:
        n = min(CHUNK_ROWS, args.rows - c * CHUNK_ROWS)
        x = gen_chunk(c, n, args.dim, device, dt)[:left]
        parts.append(x.float().cpu())
        left -= x.shape[0]
        c += 1
    X = torch.cat(parts).contiguous()
    Q = gen_queries(args.nq, args.dim, device, dt).float().cpu().numpy()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        flat_ip.flat_ip_search_fast(Q, X, args.k)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    scale = args.rows / float(sample_rows)
    return times, scale, cores, sample_rows


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    device = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    times, scale, cores, sample_rows = cpu_port_sample(args, device, args.cpu_sample_rows or 3_000_000, repeats=max(1, args.steps),
                                                       warmup=max(0, min(args.warmup, 1)))
    step_s = sum(times) / len(times) * scale
    value = args.nq / step_s
    sample = (f"all {args.nq} queries x first {sample_rows} of {args.rows} rows (fp32 upcast of the same synthetic "
              f"{args.dtype} values), {len(times)} timed passes, time scaled x{scale:.2f} to the full corpus")
    line = {
        "impl": "reference", "metric": "queries/s exact top-100 over 12Mx768 passages", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                         "sample_ms": sum(times) / len(times) * 1e3,
                         "note": "FAISS-equivalent CPU port (oracle/flat_ip.py: torch/MKL sgemm blocks + top-k merge); faiss itself is not installable offline"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"{args.rows}x{args.dim} {args.dtype} synthetic N(0,1) passage corpus, {args.nq} queries/step, exact IP top-{args.k}",
            "rows": args.rows, "dim": args.dim, "nq": args.nq, "k": args.k, "sharding": f"rows/{world}",
            "l2_policy": "inputs_larger_than_L2 (corpus shard >= 2.3 GB vs 126 MB L2)"}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from viquae_b200 import _lib
    from viquae_b200.index import B200FlatIndex
    from viquae_b200.sharded import ShardedFlatIndex, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32, "f32x3": torch.float32}[args.dtype]

    # ---- build the local shard on the device (nothing is shipped to the box)
    r0, r1 = shard_range(args.rows, rank, world)
    index = B200FlatIndex(d=args.dim, dtype=args.dtype, device=local_rank, capacity=r1 - r0, variant=args.variant, id_offset=r0)
    c = r0 // CHUNK_ROWS
    while c * CHUNK_ROWS < r1:
        c0, c1 = c * CHUNK_ROWS, min(args.rows, (c + 1) * CHUNK_ROWS)
        x = gen_chunk(c, c1 - c0, args.dim, device, dt)
        lo, hi = max(r0, c0) - c0, min(r1, c1) - c0
        index.add(x[lo:hi])
        del x
        c += 1
    assert index.ntotal == r1 - r0
    sharded = ShardedFlatIndex(index, rank=rank, world_size=world) if world > 1 else None
    q_dev = gen_queries(args.nq, args.dim, device, dt)
    q_host = q_dev.float().cpu().pin_memory()
    k = args.k
    lib = _lib.load()

    def step_device():
        if sharded is not None:
            return sharded.search_tensors(q_dev, k)
        return index.search_tensors(q_dev, k)

    out_s_host = torch.empty((args.nq, k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((args.nq, k), dtype=torch.int64).pin_memory()

    def step_e2e():
        # the call a user of the reference makes: host float32 queries in, host (scores, ids) out
        if sharded is None:
            _lib.check(lib.vq_index_search_host(index._h, q_host.data_ptr(), args.nq, 0, k, out_s_host.data_ptr(), out_i_host.data_ptr()))
        else:
            q = q_host.to(device, non_blocking=True)
            s, i = sharded.search_tensors(q, k)
            out_s_host.copy_(s, non_blocking=True)
            out_i_host.copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    # ---- device-resident throughput (+ CUDA events around the dominant kernel, recorded by the engine)
    _lib.check(lib.vq_index_set_timing(index._h, 0))
    for _ in range(args.warmup):
        step_device()
    _lib.check(lib.vq_index_set_timing(index._h, 1))
    launches0 = int(lib.vq_index_launch_count(index._h))
    total_ms, t0, t1 = timed(step_device, args.steps, 0)
    launches_timed = int(lib.vq_index_launch_count(index._h)) - launches0 + (args.steps if world > 1 else 0)
    tot = _lib.c_double()
    nl = _lib.c_int()
    _lib.check(lib.vq_index_read_timing(index._h, _lib.ctypes.byref(tot), _lib.ctypes.byref(nl)))
    _lib.check(lib.vq_index_set_timing(index._h, 0))
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- end to end through the host-buffer API
    e2e_ms, _, _ = timed(step_e2e, args.steps, args.warmup)

    latency = None
    if args.latency_calls > 0:
        # per-call latency through the host-buffer API (what interact/system.py's one-question loop would see)
        lat = []
        for _ in range(50):
            step_e2e()
        for _ in range(args.latency_calls):
            barrier()
            t_a = time.perf_counter()
            step_e2e()
            lat.append((time.perf_counter() - t_a) * 1e6)
        lat.sort()
        latency = {"calls": len(lat), "p50_us": lat[len(lat) // 2], "p99_us": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
                   "min_us": lat[0], "what": "wall clock around one host-buffer search call (H2D + search + merge + D2H + sync)"}

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = args.nq / (ms_per_step / 1e3)
        e2e_value = args.nq / (e2e_ms / args.steps / 1e3)
        peaks = {}
        peaks_src = "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peaks_src = "measured"
        except Exception:
            pass
        kern_ms = tot.value / max(1, nl.value)
        local_rows = r1 - r0
        # f32x3: six bf16 product terms per fp32 product on the tensor pipe, 12 stored bytes per element
        terms = 6 if args.dtype == "f32x3" else 1
        flops = 2.0 * args.nq * local_rows * args.dim * terms
        bytes_ = local_rows * args.dim * {"f32": 4, "f32x3": 12}.get(args.dtype, 2) + args.nq * args.dim * 2 * terms + args.nq * k * 12
        ridge_tensor = args.nq * terms > 211 and args.dtype != "f32"
        if args.dtype == "f32":
            # exact mode is SIMT FFMA-bound: 148 SMs x 128 FMA/clk x 2 flop at the max SM clock (no measured peak for it)
            peak = 148 * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
            roof = {"bound": "fp32-ffma", "achieved": flops / (kern_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s"}
        elif ridge_tensor:
            peak = float(peaks.get("bf16_tflops", 1590.0))
            roof = {"bound": "tensor", "achieved": flops / (kern_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s"}
            if terms > 1:
                roof["note"] = "bf16 tensor flops issued (6 product terms per fp32 product); fp32-equivalent rate = achieved / 6"
        else:
            peak = float(peaks.get("hbm_gbs", 6650.0))
            roof = {"bound": "hbm", "achieved": bytes_ / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["peak_source"] = f"{peaks_src} (MEASURED_PEAKS.json burst figure: kernel timed alone)" if peaks_src == "measured" else "fallback (B200_PROFILING.md)"
        # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel's launch(es), from the committed
        # `ncu --set full` capture of this same workload (profiles/traffic.json); null for other workloads
        roof["traffic"] = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tr = json.load(f)
            key = f"nq{args.nq}_rows{args.rows}_d{args.dim}_{args.dtype}_gpus{world}"
            if key in tr:
                roof["traffic"] = tr[key]["dram_bytes_per_launch"]
                roof["traffic_source"] = tr[key]["source"]
        except Exception:
            pass
        roof["kernel"] = {"f32": "f32_topk_kernel"}.get(args.dtype, "stream_topk_kernel" if (args.variant == "stream" and args.nq <= 32) else "tc_topk_kernel")
        roof["timed_region"] = "all contraction passes of a step (threshold-priming passes + main pass) and their merge kernels; excludes only the query cast"
        roof["kernel_ms"] = kern_ms
        roof["kernel_launches_timed"] = nl.value
        roof["algorithmic"] = {"flops_per_launch": flops, "bytes_per_launch": bytes_}
        line = {
            "metric": "queries/s exact top-100 over 12Mx768 passages", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": args.nq * args.dim * 4,
                    "d2h_bytes_per_step": args.nq * k * 12, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches_timed,
            "roofline": roof,
        }
        if latency is not None:
            line["latency"] = latency
        if world == 1 and not args.no_cpu_baseline:
            try:
                times, scale, cores, sample_rows = cpu_port_sample(args, device, args.cpu_sample_rows or 1_500_000, repeats=1, warmup=0)
                step_s = times[0] * scale
                line["cpu_baseline"] = {
                    "value": args.nq / step_s, "unit": "queries/s", "cores": cores, "kind": "port",
                    "sample": f"all {args.nq} queries x first {sample_rows} of {args.rows} rows (fp32 upcast of the same values), 1 pass, time scaled x{scale:.2f}",
                    "sample_ms": times[0] * 1e3}
            except Exception as e:  # the CPU leg must never sink the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
