#!/usr/bin/env python
"""bench.py -- task placements/sec on BASELINE.json's headline configuration.

Workload (config.workload): cfg3-oneoff = 1M pending one-off tasks x 100k nodes,
8 constraint expressions (node.labels / role / platform) per task, PlatformFilter
on half the services, spread ranking -- the reference's own
BenchmarkScheduler100kNodes1MTasks shape (scheduler_test.go:3358,3378-3468).

A "step" is one pass of the hot path (one Scheduler.tick group loop) over that
batch of synthetic pending tasks:
  value : placements/s with the node mirror and the tick descriptors already in
          HBM (pe_tick_run), device time from CUDA events on the engine stream
  e2e   : the same through pe_schedule with HOST buffers (pinned), H2D of the
          descriptors and D2H of the placements inside the timed region
  roofline : the scan kernel's algorithmic bytes / its CUDA-event time vs the
          measured HBM peak
  cpu_baseline : the CPU oracle (oracle/flat_oracle.cpp, 1 thread like the
          reference's single goroutine) on a bounded sample of the same workload
`--impl reference` times that CPU restatement alone (the reference is Go and no
Go toolchain exists on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "task placements/sec at 1M tasks x 100k nodes"
UNIT = "placements/s"


def scan_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per k_scan launch, from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "scan_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)["traffic_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_copy(a: np.ndarray) -> np.ndarray:
    """Copy into page-locked host memory (torch is only the allocator here)."""
    import torch
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
    v = t.numpy().view(a.dtype).reshape(a.shape)
    v[...] = a
    _KEEP.append(t)
    return v


_KEEP = []


def make_workload(name: str, n_tasks: int, n_nodes: int):
    from swarmkit_b200 import workloads as W
    base, _, mode = name.partition("-")
    kw = {}
    if base != "cfg1":
        kw = {"n_nodes": n_nodes, "n_tasks": n_tasks}
    return W.by_name(name, **kw)


def cpu_sample(w, sample_tasks: int):
    """Time the CPU oracle on the first `sample_tasks` groups of the workload (full node table)."""
    from tests.oracle_lib import OracleEngine
    o = OracleEngine(node_capacity=w.n_nodes)
    o.node_upsert(w.nodes)
    o.set_node_count(w.n_nodes)
    sub = w.tick.slice_groups(0, min(sample_tasks, w.tick.n_groups))
    t0 = time.perf_counter()
    out, _ = o.schedule(sub)
    dt = time.perf_counter() - t0
    placed = int((out != 0xFFFFFFFF).sum())
    return placed, sub.n_tasks, dt, out


def golden_prefix(args, n_tasks: int, n_nodes: int):
    """Committed full-size golden placements (tests/golden/big_*.npz, CPU oracle) for this workload, if one covers it.
    The tick is sequential, so the vector of the 1M-task tick pins every shorter tick over the same nodes."""
    table = {("cfg3-oneoff", 100_000): "big_cfg3_oneoff_1m_100k", ("cfg2-oneoff", 10_000): "big_cfg2_oneoff_100k_10k",
             ("cfg3-grouped", 100_000): "big_cfg3_grouped_1m_100k", ("cfg2-grouped", 10_000): "big_cfg2_grouped_100k_10k",
             ("cfg5-grouped", 100_000): "big_cfg5_storm_500k_100k"}
    name = table.get((args.workload, n_nodes))
    if name is None:
        return None, None
    try:
        from tests.golden import make_golden_big as GB
        gold, _ = GB.load(name)
    except Exception:
        return None, None
    if args.workload.endswith("grouped") and gold.size != n_tasks:
        return None, None       # (grouped ticks change with the replica count: only the full size is pinned)
    return (gold[:n_tasks], name) if gold.size >= n_tasks else (None, None)


def latency_arm(eng, w, reset_state, tick_tasks: int, n_ticks: int, gold):
    """BASELINE's second metric: latency of ONE schedule batch (one tick = one pe_schedule call, host buffers in, host
    buffers out) against the resident node mirror.  The pending tasks of the workload arrive `tick_tasks` at a time, the
    way the scheduler's debounce (scheduler.go:149-155) hands them over; state carries from tick to tick, and when the
    workload is used up it is offered again on top of what is already placed.  The placements of the first pass over the
    workload are, by the tick's sequential semantics, exactly the golden placements of the one big tick."""
    n = w.tick.n_groups
    per_pass = max(n // tick_tasks, 1)
    subs = [w.tick.slice_groups(i * tick_tasks, (i + 1) * tick_tasks) for i in range(per_pass)]
    reset_state()
    for s in subs[:3]:                                   # warm-up ticks (allocations, first-use paths), then a clean mirror
        eng.schedule(s)
    reset_state()
    lat, mism, checked = [], 0, 0
    for i in range(n_ticks):
        s = subs[i % per_pass]
        t0 = time.perf_counter()
        on, _ = eng.schedule(s)
        lat.append(1e3 * (time.perf_counter() - t0))
        if gold is not None and i < per_pass:
            g = gold[i * tick_tasks:(i + 1) * tick_tasks]
            mism += int((on[:g.size] != g).sum())
            checked += int(g.size)
    a = np.sort(np.array(lat))
    q = lambda p: float(a[min(len(a) - 1, int(p * len(a)))])
    return {"tasks_per_tick": tick_tasks, "ticks": n_ticks, "p50": q(0.50), "p90": q(0.90), "p99": q(0.99), "max": float(a[-1]),
            "mean": float(a.mean()), "parity": {"tasks": checked, "mismatches": mism}}


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = make_workload(args.workload, args.tasks, args.nodes)
    sample = args.cpu_sample
    vals = []
    for i in range(args.warmup + args.steps):
        placed, n, dt, _ = cpu_sample(w, sample)
        if i >= args.warmup:
            vals.append((placed, dt))
    placed = sum(p for p, _ in vals)
    tt = sum(d for _, d in vals)
    v = placed / tt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tt / max(len(vals), 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64/u32", "data": "synthetic",
        "config": {"workload": args.workload, "tasks": w.tick.n_tasks, "nodes": w.n_nodes},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"first {min(sample, w.tick.n_groups)} tasks of the workload against all {w.n_nodes} nodes, "
                                   "oracle/flat_oracle.cpp single thread (the reference scheduler is one goroutine); "
                                   "Go reference not runnable: no go toolchain"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3-oneoff")
    ap.add_argument("--tasks", type=int, default=1_000_000)
    ap.add_argument("--nodes", type=int, default=100_000)
    ap.add_argument("--cpu-sample", type=int, default=30000, help="tasks in the CPU baseline sample (~14 s on one core)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--max-batch", type=int, default=0)
    ap.add_argument("--latency-ticks", type=int, default=1000, help="ticks per size in the p99 tick-latency arm (0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from swarmkit_b200 import PlacementEngine
    from swarmkit_b200.engine import nccl_unique_id

    w = make_workload(args.workload, args.tasks, args.nodes)
    n_tasks = w.tick.n_tasks
    # N > 1: ONE scheduler whose node axis is sharded over the ranks.  Every rank mirrors all nodes and
    # submits the same tick; each scans its slice of the nodes; placements are identical on every rank.
    nccl_id = None
    if world > 1:
        box = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]
    eng = PlacementEngine(node_capacity=w.n_nodes, device=local_rank, max_batch=args.max_batch,
                          rank=rank, world_size=world, nccl_id=nccl_id)
    # pinned host copies of what crosses PCIe each step
    w.tick.groups = pinned_copy(w.tick.groups)
    w.tick.task_flags = pinned_copy(w.tick.task_flags)
    for name in ("cons", "plats", "gens", "ports"):
        a = getattr(w.tick, name)
        if a.size:
            setattr(w.tick, name, pinned_copy(a))
    # ... and of what comes back (the caller owns the result buffers of pe_schedule)
    pin_node = pinned_copy(np.zeros(max(n_tasks, 1), np.uint32))
    pin_fail = pinned_copy(np.zeros(max(w.tick.n_groups, 1) * 8, np.uint32))

    def reset_state():
        eng.node_upsert(w.nodes)   # rewrites every row and zeroes every per-service counter column
        eng.set_node_count(w.n_nodes)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: W warm-up + K timed steps
    for _ in range(args.warmup):
        reset_state()
        eng.tick_upload(w.tick)
        eng.tick_run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.stats_reset()
    dev_ms, wall_ms, placed_total = 0.0, 0.0, 0
    for _ in range(args.steps):
        reset_state()
        eng.tick_upload(w.tick)
        s0 = eng.stats()
        barrier()
        t0 = time.perf_counter()
        eng.tick_run()
        barrier()
        wall_ms += 1e3 * (time.perf_counter() - t0)
        s1 = eng.stats()
        dev_ms += s1["run_ms"] - s0["run_ms"]
        placed_total += s1["placements"] - s0["placements"]
    st = eng.stats()
    out_node, _ = eng.tick_download()
    clocks = sampler.stop()

    # ---- end-to-end arm: host buffers in, host buffers out
    e2e_ms, e2e_placed = 0.0, 0
    e2e_lat = []
    s_before = eng.stats()
    for i in range(1 + args.steps):
        reset_state()
        barrier()
        t0 = time.perf_counter()
        on, _ = eng.schedule(w.tick, pin_node, pin_fail)
        barrier()
        if i > 0:
            e2e_lat.append(1e3 * (time.perf_counter() - t0))
            e2e_ms += e2e_lat[-1]
            e2e_placed += int((on != 0xFFFFFFFF).sum())
    s_after = eng.stats()
    h2d = (s_after["h2d_bytes"] - s_before["h2d_bytes"]) // (1 + args.steps)
    d2h = (s_after["d2h_bytes"] - s_before["d2h_bytes"]) // (1 + args.steps)

    # ---- tick latency: >= 1000 ticks of 1k and of 10k pending tasks against the resident mirror
    tick_lat = {}
    if args.latency_ticks > 0 and world == 1:
        gold_l, _ = golden_prefix(args, n_tasks, w.n_nodes)
        for tt in (1000, 10000):
            if n_tasks >= tt:
                tick_lat[f"{tt // 1000}k"] = latency_arm(eng, w, reset_state, tt, args.latency_ticks, gold_l)

    # ---- max over ranks of the times, sum over ranks of the work (swarmkit_b200/dist.py)
    from swarmkit_b200.dist import reduce_step
    # the ranks hold replicas of the same decisions: count them once
    (dev_ms, wall_ms, e2e_ms), (placed_all, e2e_all) = reduce_step(
        [dev_ms, wall_ms, e2e_ms], [placed_total if rank == 0 else 0, e2e_placed if rank == 0 else 0])

    parity_failed = False
    if rank == 0:
        peak, peak_src = measured_peak()
        scan_s = st["scan_ms"] / 1e3
        achieved = st["scan_bytes"] / scan_s / 1e9 if scan_s > 0 else 0.0
        value = placed_all / (dev_ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "int64/u32", "data": "synthetic",
            "config": {"workload": args.workload, "tasks": n_tasks, "nodes": w.n_nodes, "mode": "one-off (k=1 groups)",
                       "parallelism": f"node axis sharded over {world} ranks (replicated sequencer, NCCL all-gather per batch)" if world > 1 else "1 gpu",
                       "l2": "state reset between steps rewrites every node column and 400 MB of service counters (> L2); "
                             "inside a step the 4.4 MB node table is deliberately L2-resident"},
            "wall_ms_per_step": wall_ms / args.steps,
            "placed_per_step": placed_all / args.steps,
            "evals_per_s_per_gpu": st["evals"] / scan_s if scan_s > 0 else None,
            "pairs_per_s_per_gpu": st["pairs"] / world / (dev_ms / 1e3),   # (task,node) pairs covered: each rank covers 1/N of the nodes
            "scan_rows_per_task": st["scan_rows"] / max(placed_total, 1),
            "split_ms_per_step": {"scan": st["scan_ms"] / args.steps, "place": st["place_ms"] / args.steps,
                                  "sequencer": st["sequencer_ms"] / args.steps,
                                  "classify_static_rows": st["prep_ms"] / args.steps},
            "place": {"tasks": st["place_tasks"], "handed_over_batches": st["place_cuts"], "reranked_lanes": st["place_amb"],
                      "tails": st["place_tails"], "chunks": st["place_chunks"],
                      "cta0_cycles": {"stage": st["place_cyc"][0], "resolve": st["place_cyc"][1], "commit": st["place_cyc"][2]},
                      "resolve": {"passes": st["seq_prof"][0], "rounds": st["seq_prof"][1], "short_chunks": st["seq_prof"][2],
                                  "thread0_cycles": {"filter": st["seq_prof"][3], "rounds": st["seq_prof"][4], "finalize": st["seq_prof"][5], "barrier": st["seq_prof"][6]}}},
            "paths": {"fast": st["fast_path"], "medium": st["medium_path"], "slow": st["slow_path"]},
            "sequencer_cycles": {"fast": st["seq_cycles_fast"], "medium": st["seq_cycles_medium"], "generic": st["seq_cycles_generic"],
                                 "fast_exits": st["seq_stops"], "ordered_warp_wait": st["seq_cons_wait"],
                                 "ordered_warp_work": st["seq_cons_work"], "candidate_taken": st["seq_rewalks"], "prof": st["seq_prof"]},
            "e2e": {"value": e2e_all / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    # BASELINE's second metric: latency of one schedule batch (= one tick: upload, place, download).
                    # "whole_workload": the 1M-task tick above; "1k" / "10k": latency_arm()
                    "tick_latency_ms": dict({"whole_workload": {"min": min(e2e_lat), "median": sorted(e2e_lat)[len(e2e_lat) // 2],
                                                                "max": max(e2e_lat), "samples": len(e2e_lat)}}, **tick_lat)},
            "gpu_launches": int(st["kernel_launches"]),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": scan_traffic(), "peak_source": peak_src,
                         "note": "kernel k_scan (sweep 1 + sweep 2 + k_merge per batch); algorithmic bytes per executed (row,node) "
                                 "evaluation, pre-evaluated-predicate-mask encoding of SURVEY 8d: signature bit 1/8 + total 4 + "
                                 "service count 4 [+16 cpu/mem, +8 per generic want, +4 per host-port word]; tasks with identical "
                                 "descriptors share one scan row, so evals_per_s counts executed evaluations and pairs_per_s the "
                                 "(task,node) pairs they cover; node tiles are re-used from shared memory and the node table is "
                                 "L2-resident, so the kernel is bound by integer issue, not DRAM (compare traffic with bytes_per_launch)",
                         "bytes_per_launch": st["scan_bytes"] / max(st["scan_launches"], 1),
                         "ms_per_launch": st["scan_ms"] / max(st["scan_launches"], 1)},
        }
        # ---- parity of the TIMED tick (every --gpus N): the placements it produced against (a) the committed
        # full-size golden vector of the CPU oracle and (b) the oracle run right here on the tick's first tasks
        mismatches = sum(v["parity"]["mismatches"] for v in tick_lat.values())
        gold, gold_name = golden_prefix(args, n_tasks, w.n_nodes)
        if gold is not None:
            bad = int((out_node != gold).sum())
            mismatches += bad
            line["parity_full"] = {"tasks": int(gold.size), "mismatches": bad, "golden": f"tests/golden/{gold_name}.npz"}
        if not args.no_cpu:
            placed, n, dt, cpu_out = cpu_sample(w, args.cpu_sample)
            line["cpu_baseline"] = {"value": placed / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"first {n} tasks of the same workload against all {w.n_nodes} nodes "
                                              f"({n * w.n_nodes:.3g} (task,node) visits, {dt:.1f} s), oracle/flat_oracle.cpp, 1 thread"}
            bad = int((out_node[:cpu_out.size] != cpu_out).sum())
            mismatches += bad
            line["parity_prefix"] = {"tasks": int(cpu_out.size), "mismatches": bad}
        print(json.dumps(line))
        if mismatches:
            sys.stderr.write(f"PARITY FAILURE: {mismatches} placements of the timed tick differ from the CPU oracle\n")
            parity_failed = True
    if world > 1:
        dist.destroy_process_group()
    if parity_failed:
        sys.exit(3)


if __name__ == "__main__":
    main()
