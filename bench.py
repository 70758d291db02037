#!/usr/bin/env python
"""bench.py — gate-lane evaluations per second of the B200 engine on BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 3 — synthetic random combinational DAG, 1M gates,
depth 200, fan-in 2, 4096 inputs, 2^24 stimulus lanes per GPU, 4-state engine with two-valued
stimulus generated on the device from (seed, input, global lane-word).  One step = one settle
(run_sim to the fixed point) of every lane of the batch, tile by tile (256 lane-words per tile,
two simulator replicas on two CUDA streams taking alternate tiles), plus extraction of the 1024
designated output rows.  At N GPUs every rank runs the same per-GPU batch on its own lane range
(weak scaling) and the packed output planes are all-gathered once per step with NCCL.

value      = G * L_total / max-over-ranks device time                (gate-lane evals/s)
e2e        = the same through the C ABI with HOST buffers: pinned-host stimulus planes in,
             output planes out, copies inside the timed region (overlapped with the settles)
roofline   = k_eval2 (the per-level gate kernel): 48 algorithmic bytes per gate per 64-lane
             word / its effective launch duration (timed region / launches: the
             replicas' launches overlap), against MEASURED_PEAKS.json; traffic = DRAM bytes per
             launch from the committed ncu capture (profiles/traffic.json)
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
                    help="64-lane words resident per tile and replica (0 = batch.auto_tile_words: 256 for this netlist)")
    ap.add_argument("--host-threads", type=int, default=1,
                    help="1: one host thread per replica issues its launches (0 for runs under ncu)")
    ap.add_argument("--streams", type=int, default=2, help="simulator replicas on separate CUDA streams, alternating tiles")
    ap.add_argument("--max-steps", type=int, default=10_000)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--e2e-ring-tiles", type=int, default=0, help="host stimulus ring size in tiles (0 = whole batch on 1 GPU, 2^20 lanes otherwise)")
    ap.add_argument("--cpu-lanes-per-core", type=int, default=24)
    ap.add_argument("--valid-elision", action="store_true", help="experiment: run the main timed loop with B2_OPT_VALID_ELISION (not the headline configuration)")
    ap.add_argument("--no-e2e", action="store_true")
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
        "streams": args.streams,
        "cache": "no flush: every step streams the whole wire store of each replica (GBs, far beyond the 126 MB L2) "
                 "200 levels deep; what L2 keeps between consecutive level kernels is part of the design (DESIGN.md §6)",
        "parallelism": f"lane-sharded x{n_gpus}, full netlist replica per GPU, one NCCL all-gather of output planes per step",
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
        dist.init_process_group("nccl", device_id=dev)

    from digilogic_b200 import SimulatorBuilder, SimulationRunResult
    from digilogic_b200.batch import InterleavedRunner, gather_output_planes, reduce_worst, shard_words
    from digilogic_b200.gsim import kernel_launches

    gen = build_netlist(args)
    t0 = time.perf_counter()
    first = gen.emit(SimulatorBuilder()).build()
    compile_s = time.perf_counter() - t0
    stream = torch.cuda.Stream(device=dev)   # timing events, output planes and the collective live on this stream
    torch.cuda.set_stream(stream)

    words_local = args.lanes // 64
    total_words = words_local * world
    w_begin, w_end = shard_words(total_words, world, rank)
    from digilogic_b200.batch import auto_tile_words
    tile = min(args.tile_words, words_local) if args.tile_words else auto_tile_words(first, words_local)
    args.tile_words = tile
    assert words_local % tile == 0
    made = []

    def make_sim():
        sim = first if not made else gen.emit(SimulatorBuilder()).build()
        made.append(sim)
        sim.set_device(local_rank)
        return sim

    runner = InterleavedRunner(make_sim, gen.inputs, gen.outputs, tile, args.streams, args.max_steps, device=dev,
                               host_threads=bool(args.host_threads))
    sims = runner.sims
    sim = sims[0]
    if args.valid_elision:
        for x in sims:
            x.set_option(x.OPT_VALID_ELISION, 1)
        args.no_elision_probe = True
    info = sim.info
    n_out = len(gen.outputs)
    out_v = torch.zeros((n_out, words_local), dtype=torch.int64, device=dev)
    out_m = torch.zeros_like(out_v)

    def step(record=False):
        worst = runner.run_generated(w_begin, words_local, SEED, 0, out_v, out_m, 0)
        if world > 1:
            gv, gm = gather_output_planes(out_v, out_m, world)
            worst = reduce_worst(worst, dev)
            return worst, gv, gm
        return worst, out_v, out_m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        worst, gv, gm = step()
    barrier()
    assert worst == SimulationRunResult.Ok

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        worst, gv, gm = step(record=True)
    e1.record(stream)
    barrier()
    launches = kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # a checksum of the outputs, so that the timed work is observable (and comparable across N)
    checksum = int((gv.to(torch.int64).sum() ^ gm.to(torch.int64).sum()).item()) & 0xFFFFFFFFFFFFFFFF

    gate_lane_evals = float(args.gates) * args.lanes * world * args.steps
    value = gate_lane_evals / (ms * 1e-3)

    # roofline of the dominant kernel (k_eval2, one launch per level per tile).  With several replicas in
    # flight the launches overlap, so the per-launch duration is the effective one: timed region / launches
    # (the few other kernels of a step, 0.6 % of it, are charged to k_eval2).
    n_tiles = words_local // tile
    eval_launches = n_tiles * info.depth * args.steps
    avg_launch_s = ms * 1e-3 / eval_launches
    gate_words_per_launch = (args.gates / info.depth) * info.row_stride
    algo_bytes_per_launch = ALGO_BYTES_PER_GATE_WORD * gate_words_per_launch
    peak, peak_kind = load_peaks()
    achieved = algo_bytes_per_launch / avg_launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "kernel": "k_eval2<4,false,false>",
                "algorithmic_bytes_per_launch": algo_bytes_per_launch, "avg_launch_us": avg_launch_s * 1e6,
                "launches_timed": eval_launches,
                "note": "effective launch duration (timed region / launches; replicas overlap). frac > 1: a level's output rows are "
                        "still in L2 when the next level reads them, so about a third of the algorithmic fan-in bytes never reach HBM "
                        "(see traffic)"}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            with open(traffic_path) as f:
                tj = json.load(f)
            if tj.get("tile_words") == tile and tj.get("gates") == args.gates and tj.get("streams", 1) == args.streams:
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
        except Exception:
            pass

    # secondary figure, not the headline: the same job with valid-plane elision (rows that are provably
    # valid on every lane keep no valid plane in HBM: 24 B instead of 48 B per gate-word; bit-identical results)
    elision = None
    if not args.no_elision_probe:
        for x in sims:
            x.set_option(x.OPT_VALID_ELISION, 1)
        worst, gv2, gm2 = step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        n_el = max(1, min(args.steps, 2))
        for _ in range(n_el):
            worst, gv2, gm2 = step()
        f1.record(stream)
        barrier()
        ms2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        ms2 = float(ms2.item())
        cs2 = int((gv2.to(torch.int64).sum() ^ gm2.to(torch.int64).sum()).item()) & 0xFFFFFFFFFFFFFFFF
        elision = {"value": float(args.gates) * args.lanes * world * n_el / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / n_el,
                   "steps": n_el, "bytes_per_gate_word": 24, "output_checksum_equal": cs2 == checksum,
                   "note": "B2_OPT_VALID_ELISION=1; all stimulus lanes valid, so no valid plane is read or written"}
        for x in sims:
            x.set_option(x.OPT_VALID_ELISION, 0)
        step()   # back to plain two-plane rows before the end-to-end leg
        barrier()

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, sim, runner, gen, world, rank, dev, stream, words_local, w_begin, barrier,
                      (out_v, out_m))

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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 bit-planes (value+valid)", "data": "synthetic",
            "config": config_dict(args, world, {"tiles_per_step": n_tiles, "row_stride_words": info.row_stride,
                                                "store_bytes": info.store_bytes * len(sims), "compile_s": round(compile_s, 3)}),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "valid_elision": elision, "unit_delay": unit_delay, "gpu_launches": int(launches), "clocks": clocks,
            "output_checksum": f"{checksum:#018x}",
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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


def run_e2e(args, torch, dist, sim, runner, gen, world, rank, dev, stream, words_local, w_begin, barrier, dev_out):
    """Same metric through the C ABI with host buffers (b2_set_drives_lanes / b2_gather_wires_host_async):
    every tile's stimulus planes are copied from pinned host memory and its output planes back to
    pinned host memory inside the timed region."""
    n_in, n_out = len(gen.inputs), len(gen.outputs)
    tile = runner.tile_words
    # One GPU: host planes for the whole batch (17.2 + 4.3 GB).  Several GPUs share one host: a ring of
    # 2^20 lanes (1.3 GB) per rank bounds the pinned memory; every tile is still copied in and out.
    ring_tiles = args.e2e_ring_tiles or (words_local // tile if world == 1 else max(8, 16384 // tile))
    ring_tiles = min(ring_tiles, words_local // tile)
    ring_words = ring_tiles * tile
    try:
        hin_v = torch.empty((n_in, ring_words), dtype=torch.int64, pin_memory=True)
        hin_m = torch.empty((n_in, ring_words), dtype=torch.int64, pin_memory=True)
        hout_v = torch.empty((n_out, ring_words), dtype=torch.int64, pin_memory=True)
        hout_m = torch.empty((n_out, ring_words), dtype=torch.int64, pin_memory=True)
    except Exception as ex:  # not enough pinnable host memory on this box
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}
    # host stimulus = the same counter-based words, produced once on the device and copied out before
    # the timed region (numpy generation of 17 GB would take minutes): after a settle the source rows
    # of the store hold the drives, so gathering the input rows returns the stimulus planes
    torch.cuda.synchronize()
    for w0 in range(0, ring_words, tile):
        sim.fill_stimulus(gen.inputs, SEED, lane_word_base=w_begin + w0, mode=0)
        sim.run_sim_lanes_async(args.max_steps)
        sim.gather_wires_ptr(gen.inputs, hin_v.data_ptr() + 8 * w0, hin_m.data_ptr() + 8 * w0, ring_words, device=False)
    sim.synchronize()
    inv, inm, outv, outm = hin_v.numpy(), hin_m.numpy(), hout_v.numpy(), hout_m.numpy()
    worst = runner.run_host(inv, inm, outv, outm, words_local)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        if world > 1:
            # host planes go to this rank's host memory; the cross-rank gather moves the device copies
            from digilogic_b200.batch import gather_output_planes
            worst = runner.run_host(inv, inm, outv, outm, words_local, dev_out[0], dev_out[1])
            gather_output_planes(dev_out[0], dev_out[1], world)
        else:
            worst = runner.run_host(inv, inm, outv, outm, words_local)   # returns after the last D2H copy
    barrier()
    ms = (time.perf_counter() - t0) * 1e3   # host clock: the path ends in host memory
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = float(args.gates) * args.lanes * world * args.e2e_steps / (ms * 1e-3)
    return {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(n_in * words_local * 16),
            "d2h_bytes_per_step": int(n_out * words_local * 16), "ms_per_step": ms / args.e2e_steps, "steps": args.e2e_steps,
            "result": int(worst), "host_ring_tiles": ring_tiles, "tiles_per_step": words_local // tile,
            "overlap": "H2D of a replica's next tile and D2H of its previous one overlap the settles (engine copy streams)",
            "host_checksum": f"{int(outv.sum() ^ outm.sum()) & 0xFFFFFFFFFFFFFFFF:#018x}"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
