#!/usr/bin/env python
"""bench.py — gate-lane evaluations per second of the B200 engine on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 3 — synthetic random combinational DAG, 1M gates,
depth 200, fan-in 2, 4096 inputs, 2^24 stimulus lanes per GPU, 4-state engine with two-valued
stimulus generated on the device from (seed, input, global lane-word).  One step = one settle
(run_sim to the fixed point) of every lane of the batch plus extraction of the 1024 designated
output rows = ONE call of the C ABI's batch runner (b2_batch_run_generated): tiles of 256 lane-words
on three simulator replicas that share one device copy of the netlist, each replaying its level sweep
from a CUDA graph on its own stream (--sweep: the persistent level-sweep kernel instead, one launch
per step).  At N GPUs every rank runs the same per-GPU batch on its own lane range (weak scaling); the
output rows of every tile are stored into every rank's gathered planes over NVLink by the kernel that
reads them (CUDA IPC mappings; --nccl-gather: by ncclAllGather after the last tile).

value      = G * L_total / max-over-ranks device time                (gate-lane evals/s)
e2e        = the same through the C ABI with HOST buffers (b2_batch_run_planes): pinned-host stimulus
             planes in, output planes out, both moved inside the timed region every step
roofline   = k_eval2<2,false,false>, the per-level gate kernel: 48 algorithmic bytes per gate per 64-lane
             word / its effective launch duration (timed region / launches: the replicas'
             launches overlap), against MEASURED_PEAKS.json; traffic / dram_frac = DRAM bytes of one whole
             step with all replicas running (ncu range replay, profiles/traffic.json) per launch
shard_check (N > 1) = rank 0 settles a tile of another rank's lanes itself and compares with the gathered planes
config5    (N > 1) = secondary: BASELINE config 5 (10M gates, 2^26 lanes sharded over the N GPUs)
cpu_baseline = the scalar event-driven gsim restatement (oracle/gsim_oracle.c, "port") on the
             host cores over a bounded lane sample of the same netlist
valid_elision = secondary figure (never the headline): the same job with B2_OPT_VALID_ELISION
unit_delay = secondary figure: BASELINE config 4 (flip-flop LFSR chains + SR latches, feedback) on the
             unit-delay path — 65 536 lanes, 20 clocked steps = 40 run_sim(10_000) calls, wall time of the calls
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

METRIC = "gate_lane_evals_per_s"
UNIT = "gate-lane evals/s"
SEED = 0xD1610003
ALGO_BYTES_PER_GATE_WORD = 48.0   # 2 fan-ins x 2 planes x 8 B read + 2 planes x 8 B written


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gates", type=int, default=1_000_000)
    ap.add_argument("--depth", type=int, default=200)
    ap.add_argument("--inputs", type=int, default=4096)
    ap.add_argument("--outputs", type=int, default=1024)
    ap.add_argument("--lanes", type=int, default=1 << 24, help="stimulus lanes per GPU")
    ap.add_argument("--tile-words", type=int, default=0,
                    help="64-lane words per tile and slot (0 = batch.auto_batch_shape: 256 for this netlist)")
    ap.add_argument("--slots", type=int, default=0, help="tiles in flight = simulator replicas (0 = batch.auto_batch_shape: 3)")
    ap.add_argument("--sweep", action="store_true", help="use the persistent level-sweep kernel (B2_BATCH_OPT_SWEEP) instead of the replicas")
    ap.add_argument("--task-iters", type=int, default=1, help="256-thread iterations per gate task of the sweep")
    ap.add_argument("--unroll", type=int, default=4, help="gates per thread of the one/two-input class in the sweep (2 or 4)")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="resident CTAs per SM of the sweep (0 = as many as fit)")
    ap.add_argument("--nccl-gather", action="store_true",
                    help="N > 1: move the output planes with ncclAllGather after the sweep instead of storing them into the peers' "
                         "planes from the sweep kernel (the baseline the fused gather is compared with)")
    ap.add_argument("--no-config5", action="store_true", help="N > 1: skip the secondary config-5 settle")
    ap.add_argument("--max-steps", type=int, default=10_000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-ring-tiles", type=int, default=0, help="host stimulus ring size in tiles (0 = whole batch on 1 GPU, 2^20 lanes otherwise)")
    ap.add_argument("--cpu-lanes-per-core", type=int, default=24)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pair-fusion", action="store_true", help="B2_OPT_PAIR_FUSION = 0: one launch per level (the library default is automatic: fused for config 3, per level for config 5)")
    ap.add_argument("--no-elision-probe", action="store_true", help="skip the secondary run with B2_OPT_VALID_ELISION")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-unit-delay-probe", action="store_true", help="skip the secondary config-4 (feedback) run")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_netlist(args):
    from digilogic_b200 import netgen
    return netgen.random_dag(n_gates=args.gates, depth=args.depth, n_inputs=args.inputs, n_outputs=args.outputs, seed=SEED)


def config_dict(args, n_gpus, extra=None):
    cfg = {
        "workload": f"BASELINE config 3: random combinational DAG, {args.gates} gates, depth {args.depth}, fan-in 2, "
                    f"{args.inputs} inputs, {args.lanes} lanes per GPU (4-state engine, two-valued device-generated stimulus)",
        "gates": args.gates, "depth": args.depth, "inputs": args.inputs, "outputs": args.outputs,
        "lanes_per_gpu": args.lanes, "lanes_total": args.lanes * n_gpus, "tile_words": args.tile_words,
        "max_steps": args.max_steps, "seed": hex(SEED),
        "cache": "no flush: every step streams the whole wire store of each slot (GBs, far beyond the 126 MB L2) "
                 "200 levels deep; what L2 keeps between consecutive levels is part of the design (DESIGN.md §6)",
        "parallelism": f"lane-sharded x{n_gpus}, full netlist replica per GPU, output planes gathered into every rank each step",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------- CPU arms

def cpu_port_run(gen, n_lanes, threads, max_steps):
    """Scalar event-driven gsim restatement (oracle/, the checker) on `n_lanes` lanes of the
    workload.  Returns (seconds, gate-lane evals/s)."""
    import oracle
    from digilogic_b200 import netgen
    ob = gen.emit(oracle.Builder())
    lanes = oracle.Lanes(ob, n_lanes)
    words = (n_lanes + 63) // 64
    v, m = netgen.stimulus_words(SEED, len(gen.inputs), words, 0, 0)
    for i, w in enumerate(gen.inputs):
        lanes.set_drive(int(w), v[i], m[i])
    t0 = time.perf_counter()
    status, _, _ = lanes.run(max_steps, threads=threads)
    dt = time.perf_counter() - t0
    assert (status == 0).all()
    return dt, gen.n_gates * n_lanes / dt


def cpu_bitsliced_run(gen, threads, words_per_thread=8):
    """The "best CPU" context figure of SURVEY.md section 8d: the dense 64-lanes-per-word model of the same semantics
    (oracle/bitsliced.c), one instance per host thread over disjoint lanes, one levelised pass each.  Not gsim's
    algorithm — what a CPU can do with the engine's own data layout.  Returns (seconds, gate-lane evals/s)."""
    import oracle
    from digilogic_b200 import netgen
    ob = gen.emit(oracle.Builder())
    n_lanes = 64 * words_per_thread
    sims = [oracle.BitSliced(ob, n_lanes, gen.n_wires) for _ in range(threads)]
    for t, bs in enumerate(sims):
        v, m = netgen.stimulus_words(SEED, len(gen.inputs), words_per_thread, t * words_per_thread, 0)
        for i, w in enumerate(gen.inputs):
            bs.set_drive(int(w), v[i], m[i])
    errs = []

    def work(bs):
        if bs.eval_dag().any():
            errs.append("conflict")

    th = [threading.Thread(target=work, args=(bs,)) for bs in sims]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    assert not errs
    return dt, gen.n_gates * n_lanes * threads / dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  gsim 1.1.4 itself cannot
    be built here (Rust crate, not vendored, no cargo), so this is the oracle port on all host
    cores, on a bounded lane sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    gen = build_netlist(args)
    n_lanes = max(cores * max(args.cpu_lanes_per_core // 2, 1), 1)
    times = []
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue  # one warm-up pass is enough to fault the netlist in; CPU passes take seconds
        dt, _ = cpu_port_run(gen, n_lanes, cores, args.max_steps)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = gen.n_gates * n_lanes * len(times) / total
    sample = f"{n_lanes} lanes of the {args.lanes}-lane batch per step, {cores} threads (one simulator per lane)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 bit-planes (CPU: u32 atoms)", "data": "synthetic",
        "config": config_dict(args, args.gpus, {"note": "CPU restatement of gsim 1.1.4 (oracle port), not the Rust crate"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL honours NCCL_DEBUG_FILE only above the VERSION level
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)   # the harness's channel for two 128-byte blobs and the timing reduction

    from digilogic_b200 import SimulatorBuilder, SimulationRunResult
    from digilogic_b200.batch import Batch, auto_batch_shape, setup_comm, shard_words
    from digilogic_b200.gsim import kernel_launches

    gen = build_netlist(args)
    t0 = time.perf_counter()
    sim = gen.emit(SimulatorBuilder()).build()            # compiled once; every slot shares the one device copy
    compile_s = time.perf_counter() - t0
    stream = torch.cuda.Stream(device=dev)   # the batch runs on this stream; the timing events are recorded on it
    torch.cuda.set_stream(stream)

    words_local = args.lanes // 64
    total_words = words_local * world
    w_begin, w_end = shard_words(total_words, world, rank)
    auto_tile, auto_slots = auto_batch_shape(sim, words_local)
    tile = min(args.tile_words, words_local) if args.tile_words else auto_tile
    slots = args.slots or auto_slots
    args.tile_words, args.slots = tile, slots
    if args.no_pair_fusion:
        sim.set_option(sim.OPT_PAIR_FUSION, 0)
    batch = Batch(sim, gen.inputs, gen.outputs, tile, slots, device=local_rank, stream=stream.cuda_stream)
    batch.set_option(Batch.OPT_SWEEP, 1 if args.sweep else 0)
    batch.set_option(Batch.OPT_TASK_ITERS, args.task_iters)
    batch.set_option(Batch.OPT_UNROLL, args.unroll)
    if args.ctas_per_sm:
        batch.set_option(Batch.OPT_CTAS_PER_SM, args.ctas_per_sm)
    info = sim.info
    n_out = len(gen.outputs)
    out_v = out_m = None
    if world > 1:
        # gathered planes [world, n_out, words_local] on every rank; with the peers' planes mapped (CUDA IPC) the kernel that
        # reads a tile's output rows stores them into all of them over NVLink
        setup_comm(batch, words_local, local_rank, map_peers=not args.nccl_gather)
    else:
        out_v = torch.zeros((n_out, words_local), dtype=torch.int64, device=dev)
        out_m = torch.zeros_like(out_v)

    def step():
        # ONE C call per step: stimulus, settle, output rows to their destinations, (N > 1) the ranks meet
        batch.run_generated(SEED, 0, w_begin, words_local, args.max_steps, out_v, out_m)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def planes():
        if world > 1:
            return tuple(torch.as_tensor(x, device=dev) for x in batch.gathered())
        return out_v, out_m

    for _ in range(args.warmup):
        step()
    worst = batch.wait()
    barrier()
    assert worst == SimulationRunResult.Ok
    binfo = batch.info
    assert binfo.sweep == (1 if args.sweep else 0)

    pair_plan = None
    if binfo.pair_launches:
        pair_plan = sim.plan_pairs(gen.outputs)   # host-side statistics of the plan the replicas run (same planner, same kept rows)
        assert pair_plan["launches"] == binfo.pair_launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    worst = batch.wait()
    barrier()
    launches = kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    assert worst == SimulationRunResult.Ok
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # a checksum of the outputs, so that the timed work is observable (and comparable across N)
    gv, gm = planes()
    checksum = int((gv.sum() ^ gm.sum()).item()) & 0xFFFFFFFFFFFFFFFF

    # N > 1: rank 0 settles one tile of ANOTHER rank's lane range itself and compares it with what that rank gathered
    shard_check = None
    if world > 1:
        other = world - 1
        chk_words = min(tile, words_local)
        ob, _ = shard_words(total_words, world, other)
        probe = Batch(sim, gen.inputs, gen.outputs, chk_words, 1, device=local_rank, stream=stream.cuda_stream)
        pv = torch.zeros((n_out, chk_words), dtype=torch.int64, device=dev)
        pm = torch.zeros_like(pv)
        probe.run_generated(SEED, 0, ob + words_local - chk_words, chk_words, args.max_steps, pv, pm)   # the other rank's last tile
        probe.wait()
        ok = torch.equal(gv[other][:, words_local - chk_words:], pv) and torch.equal(gm[other][:, words_local - chk_words:], pm)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)     # every rank checks its copy of the gathered planes
        shard_check = bool(flag.item())
        del probe

    gate_lane_evals = float(args.gates) * args.lanes * world * args.steps
    value = gate_lane_evals / (ms * 1e-3)

    # roofline of the dominant kernel.  Replicas: k_eval2, one launch per level per tile; the replicas' launches overlap, so
    # the per-launch duration is the effective one, timed region / launches (the few other kernels of a tile — stimulus,
    # drives, output rows: 0.6 % of a step — are charged to it).  --sweep: k_sweep, one launch per step.
    # Algorithmic bytes: 48 B per gate per 64-lane word.
    n_tiles = -(-words_local // tile)
    gate_words = float(args.gates) * words_local
    step_s = ms * 1e-3 / args.steps
    if args.sweep:
        kernel, launches_per_step = f"k_sweep<{args.unroll}>", 1
    else:
        per_level_ctas4 = -(-(args.gates // info.depth) // (max(256 // max(tile // 2, 1), 1) * 4)) * max((tile // 2 + 255) // 256, 1)
        kernel = "k_eval2<4,false,false>" if per_level_ctas4 >= 148 * 3 * 4 else "k_eval2<2,false,false>"   # Engine::run_levelised's rule
        launches_per_step = n_tiles * info.depth
        if binfo.pair_launches:      # level-pair fusion (B2_OPT_PAIR_FUSION): items of two dependent gates, about depth / 2 launches per tile
            kernel, launches_per_step = "k_eval_pair", n_tiles * binfo.pair_launches
    launch_s = step_s / launches_per_step
    algo_bytes = ALGO_BYTES_PER_GATE_WORD * gate_words / launches_per_step
    peak, peak_kind = load_peaks()
    achieved = algo_bytes / launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "dram_frac": None, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                "kernel": kernel, "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_us": launch_s * 1e6,
                "launches_timed": launches_per_step * args.steps,
                "note": "effective launch duration (timed region / launches; the replicas overlap). frac is the ALGORITHMIC fraction: "
                        "> 1 because a level's output rows are still in L2 when the next level reads them. dram_frac = measured DRAM "
                        "bytes (traffic: one whole step with all replicas running, ncu range replay, per launch) / duration / peak "
                        "is the physical fraction"}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            with open(traffic_path) as f:
                tj = json.load(f)
            if (tj.get("engine") == ("sweep" if args.sweep else "replicas") and tj.get("tile_words") == tile
                    and tj.get("slots") == slots and tj.get("gates") == args.gates
                    and bool(tj.get("pair_fusion")) == bool(binfo.pair_launches)):
                roofline["traffic"] = tj["dram_bytes_per_gate_word"] * gate_words / launches_per_step
                roofline["dram_frac"] = roofline["traffic"] / launch_s / 1e9 / peak
                roofline["l2_hit_rate_pct"] = tj.get("l2_hit_rate_pct")
        except Exception:
            pass

    config5 = None
    if world > 1 and not args.no_config5:
        config5 = config5_probe(args, torch, dist, world, rank, local_rank, dev, stream)

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, batch, sim, gen, world, rank, dev, words_local, barrier)

    # secondary figure, not the headline: B2_OPT_VALID_ELISION on the simulator-replica path of the batch runner (rows that are
    # provably valid on every lane keep no valid plane in HBM: 24 B instead of 48 B per gate-word; bit-identical results)
    elision = None
    if world == 1 and not args.no_elision_probe:
        sim.set_option(sim.OPT_VALID_ELISION, 1)
        eb = Batch(sim, gen.inputs, gen.outputs, tile, slots, device=local_rank, stream=stream.cuda_stream)
        ev2, em2 = torch.zeros_like(out_v), torch.zeros_like(out_m)
        eb.run_generated(SEED, 0, w_begin, words_local, args.max_steps, ev2, em2)
        eb.wait()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        eb.run_generated(SEED, 0, w_begin, words_local, args.max_steps, ev2, em2)
        f1.record(stream)
        eb.wait()
        ms2 = f0.elapsed_time(f1)
        cs2 = int((ev2.sum() ^ em2.sum()).item()) & 0xFFFFFFFFFFFFFFFF
        elision = {"value": float(args.gates) * args.lanes / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "steps": 1,
                   "bytes_per_gate_word": 24, "output_checksum_equal": cs2 == checksum,
                   "note": "B2_OPT_VALID_ELISION=1 on the per-level kernels (same batch shape); "
                           "all stimulus lanes valid, so no valid plane is read or written"}
        sim.set_option(sim.OPT_VALID_ELISION, 0)
        del eb, ev2, em2

    unit_delay = None
    if rank == 0 and world == 1 and not args.no_unit_delay_probe:
        unit_delay = unit_delay_probe(args)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_cpu_lanes = cores * args.cpu_lanes_per_core
        dt, v = cpu_port_run(gen, n_cpu_lanes, cores, args.max_steps)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu_lanes} of the {args.lanes} lanes, {dt:.1f} s, scalar event-driven gsim restatement "
                         f"(oracle/gsim_oracle.c), one simulator per lane, {cores} threads"}
        dtb, vb = cpu_bitsliced_run(gen, cores)
        cpu["bitsliced"] = {"value": vb, "unit": UNIT, "cores": cores,
                            "sample": f"{cores * 512} lanes, {dtb:.2f} s, dense 64-lane bit-sliced CPU model (oracle/bitsliced.c), one "
                                      f"levelised pass per thread: a best-CPU context figure, not gsim's algorithm"}
        cpu["gpu_over_scalar_port"] = value / v
        cpu["gpu_over_bitsliced_cpu"] = value / vb

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 bit-planes (value+valid)", "data": "synthetic",
            "config": config_dict(args, world, {"tiles_per_step": n_tiles, "slots": slots, "row_stride_words": tile,
                                                "store_bytes": binfo.store_bytes, "compile_s": round(compile_s, 3),
                                                "grid_ctas": binfo.grid_ctas, "gates_per_task": binfo.gates_per_task,
                                                "tasks_per_tile": binfo.tasks_per_pass,
                                                "gather": None if world == 1 else ("ncclAllGather" if args.nccl_gather else
                                                                                   "fused: the kernel that reads a tile's output rows stores them into every rank's planes over NVLink (CUDA IPC)"),
                                                "engine": "sweep" if args.sweep else "replicas",
                                                "pair_fusion": pair_plan}),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "valid_elision": elision, "unit_delay": unit_delay,
            "gpu_launches": int(launches), "clocks": clocks, "output_checksum": f"{checksum:#018x}",
        }
        if shard_check is not None:
            line["shard_check"] = shard_check
        if config5 is not None:
            line["config5"] = config5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def config5_probe(args, torch, dist, world, rank, local_rank, dev, stream, total_lanes=1 << 26):
    """Secondary key at N > 1: BASELINE config 5 — 10M-gate DAG, 2^26 lanes sharded contiguously over the N GPUs, output
    planes gathered into every rank — one warm-up on an eighth of the shard, then one timed settle of the whole shard."""
    from digilogic_b200 import SimulatorBuilder, netgen
    from digilogic_b200.batch import Batch, auto_batch_shape, setup_comm, shard_words
    gates, depth, n_in, n_out = 10_000_000, 200, 16384, 1024
    t0 = time.perf_counter()
    gen = netgen.random_dag(n_gates=gates, depth=depth, n_inputs=n_in, n_outputs=n_out, seed=0xD1610005)
    sim = gen.emit(SimulatorBuilder()).build()
    build_s = time.perf_counter() - t0
    words_local = total_lanes // 64 // world
    w_begin, _ = shard_words(words_local * world, world, rank)
    tile, slots = auto_batch_shape(sim, words_local)
    batch = Batch(sim, gen.inputs, gen.outputs, tile, slots, device=local_rank, stream=stream.cuda_stream)
    setup_comm(batch, words_local, local_rank, map_peers=not args.nccl_gather)
    batch.run_generated(0xD1610005, 0, w_begin, max(words_local // 8, tile), args.max_steps)
    batch.wait()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    batch.run_generated(0xD1610005, 0, w_begin, words_local, args.max_steps)
    e1.record(stream)
    worst = batch.wait()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    gv, gm = (torch.as_tensor(x, device=dev) for x in batch.gathered())
    checksum = int((gv.sum() ^ gm.sum()).item()) & 0xFFFFFFFFFFFFFFFF
    peak, _ = load_peaks()
    value = float(gates) * total_lanes / (ms * 1e-3)
    out = {"workload": f"BASELINE config 5: random DAG, {gates} gates, depth {depth}, {n_in} inputs, {total_lanes} lanes sharded over "
                       f"{world} GPUs ({words_local * 64} per GPU), {n_out} output rows gathered into every rank",
           "value": value, "unit": UNIT, "ms": ms, "steps": 1, "result": int(worst), "tile_words": tile, "slots": slots,
           "roofline_frac_per_gpu": ALGO_BYTES_PER_GATE_WORD * gates * words_local / (ms * 1e-3) / 1e9 / peak,
           "netgen_and_compile_s": round(build_s, 1), "store_bytes": batch.info.store_bytes,
           "output_checksum": f"{checksum:#018x}"}
    del batch
    return out


def unit_delay_probe(args, n_lanes=65536, clocks=20):
    """BASELINE config 4 at full netlist size on the unit-delay path (feedback: latches and flip-flops), a short
    clocked run; every lane's LFSR state is checked against a software LFSR afterwards."""
    import time
    import numpy as np
    from digilogic_b200 import SimulationRunResult, SimulatorBuilder, netgen
    bits = chains = 64
    seq = netgen.sequential(n_latches=4096, n_chains=chains, bits=bits, n_hazard=0)
    taps = [bits - 1, bits - 2, bits - 4, bits - 5]
    sim = seq.emit(SimulatorBuilder()).build()
    sim.set_lanes(n_lanes)
    words = n_lanes // 64
    all1 = np.uint64(0xFFFFFFFFFFFFFFFF)
    ones, zeros = np.full(words, all1, np.uint64), np.zeros(words, np.uint64)
    rng = np.random.default_rng(4)
    patterns = [rng.integers(0, 1 << 64, words, dtype=np.uint64) for _ in range(bits)]
    sim.set_wire_drive_lanes(seq.clk, zeros, ones)
    for b in range(bits):
        sim.set_wire_drive_lanes(seq.preset_n[b], ~patterns[b], ones)
        sim.set_wire_drive_lanes(seq.clear_n[b], patterns[b], ones)
    sim.run_sim_lanes(args.max_steps, masks=False)
    for b in range(bits):
        sim.set_wire_drive_lanes(seq.preset_n[b], ones, ones)
        sim.set_wire_drive_lanes(seq.clear_n[b], ones, ones)
    sim.run_sim_lanes(args.max_steps, masks=False)
    sim.synchronize()
    apps, t = 0, 0.0
    for _ in range(clocks):
        for level in (ones, zeros):
            sim.set_wire_drive_lanes(seq.clk, level, ones)
            t0 = time.perf_counter()
            res, _, _ = sim.run_sim_lanes(args.max_steps, masks=False)   # returns when the run has settled on the device
            t += time.perf_counter() - t0
            assert res == SimulationRunResult.Ok
            apps += sim.info.last_applications
    # lane-word 0 of chain 0 against a software Fibonacci LFSR
    q = np.stack([np.unpackbits(sim.get_wire_lanes(seq.q[0][b])[0][:1].view(np.uint8), bitorder="little") for b in range(bits)])
    start = np.stack([np.unpackbits(patterns[b][:1].view(np.uint8), bitorder="little") for b in range(bits)])
    ok = True
    for lane in range(64):
        st = sum(int(start[b][lane]) << b for b in range(bits))
        for _ in range(clocks):
            fb = 0
            for tp in taps:
                fb ^= (st >> tp) & 1
            st = ((st << 1) | fb) & ((1 << bits) - 1)
        ok &= st == sum(int(q[b][lane]) << b for b in range(bits))
    runs = 2 * clocks
    return {"workload": f"BASELINE config 4: {chains} LFSR chains x {bits} NAND flip-flops + 4096 SR latches ({seq.n_gates} gates, feedback), "
                        f"{n_lanes} lanes, {clocks} clocked steps = {runs} run_sim({args.max_steps}) calls",
            "value": seq.n_gates * n_lanes * runs / t, "unit": UNIT, "ms_per_run_sim": 1e3 * t / runs,
            "gate_lane_steps_per_s": seq.n_gates * n_lanes * apps / t, "applications": int(apps),
            "lfsr_state_matches_software": bool(ok)}


def run_e2e(args, torch, dist, batch, batch_sim, gen, world, rank, dev, words_local, barrier):
    """Same metric through the C ABI with HOST buffers: one b2_batch_run_planes call per step copies every tile's stimulus planes
    from page-locked host memory and its output planes back to page-locked host memory — both inside the timed region, both
    every step, on the replicas' copy streams while other tiles settle."""
    n_in, n_out = len(gen.inputs), len(gen.outputs)
    tile = args.tile_words
    # One GPU: host planes for the whole batch (17.2 + 4.3 GB).  Several GPUs share one host: a ring of a quarter of the
    # shard (2^22 lanes, 5.4 GB per rank) bounds the pinned memory; every tile is still read and written over PCIe.
    ring_words = args.e2e_ring_tiles * tile if args.e2e_ring_tiles else (words_local if world == 1 else max(words_local // 4, tile))
    ring_words = min(ring_words, words_local)
    try:
        hin_v = torch.empty((n_in, ring_words), dtype=torch.int64, pin_memory=True)
        hin_m = torch.empty((n_in, ring_words), dtype=torch.int64, pin_memory=True)
        hout_v = torch.zeros((n_out, ring_words), dtype=torch.int64, pin_memory=True)
        hout_m = torch.zeros((n_out, ring_words), dtype=torch.int64, pin_memory=True)
    except Exception as ex:  # not enough pinnable host memory on this box
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}
    # host stimulus = the same counter-based words, produced once on the device before the timed region (numpy generation of
    # 17 GB would take minutes): a helper batch whose "outputs" are the input wires stores them straight into the pinned planes
    from digilogic_b200.batch import Batch
    w_begin = rank * words_local
    helper = Batch(batch_sim, gen.inputs, gen.inputs, tile, args.slots, device=dev.index)
    helper.run_generated(SEED, 0, w_begin, ring_words, args.max_steps, hin_v, hin_m)
    helper.wait()
    del helper
    ring = ring_words if ring_words < words_local else 0

    def step():
        batch.run_planes(hin_v, hin_m, words_local, args.max_steps, hout_v, hout_m, in_ring_words=ring, out_ring_words=ring)

    step()
    worst = batch.wait()      # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step()
    worst = max(worst, batch.wait())   # returns when the last output row is in host memory (and, N > 1, the ranks have met)
    barrier()
    ms = (time.perf_counter() - t0) * 1e3   # host clock: the path ends in host memory
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = float(args.gates) * args.lanes * world * args.e2e_steps / (ms * 1e-3)
    outv, outm = hout_v.numpy(), hout_m.numpy()
    return {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(n_in * words_local * 16),
            "d2h_bytes_per_step": int(n_out * words_local * 16), "ms_per_step": ms / args.e2e_steps, "steps": args.e2e_steps,
            "result": int(worst), "host_ring_words": ring_words, "tiles_per_step": -(-words_local // tile),
            "path": "b2_batch_run_planes, one C call per step: H2D of a tile's stimulus planes and D2H of its output planes on the "
                    "replicas' copy streams, overlapped with the settles of the other tiles",
            "host_checksum": f"{int(outv.sum() ^ outm.sum()) & 0xFFFFFFFFFFFFFFFF:#018x}"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
