#!/usr/bin/env python
"""Shape sweep of the level-sweep kernel on BASELINE config 3: tile_words x slots x task iterations x unroll x CTAs/SM.
One JSON line per shape (ms per settle of --lanes lanes, gate-lane evals/s, algorithmic fraction of the measured HBM peak).

    python benchmarks/sweep_probe.py --lanes 4194304 --shapes 64x4x1x4x0,128x4x1x4x0,...
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gates", type=int, default=1_000_000)
    ap.add_argument("--depth", type=int, default=200)
    ap.add_argument("--inputs", type=int, default=4096)
    ap.add_argument("--outputs", type=int, default=1024)
    ap.add_argument("--lanes", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0xD1610003)
    ap.add_argument("--eval-unroll", type=int, default=0, help="B2_OPT_EVAL_UNROLL on the replicas (0 = automatic, 2, 4)")
    ap.add_argument("--elision", type=int, default=0, help="B2_OPT_VALID_ELISION on the replicas")
    ap.add_argument("--shapes", default="64x4x1x4x0,64x8x1x4x0,128x4x1x4x0,128x2x1x4x0,256x2x1x4x0,256x3x1x4x0,32x8x1x4x0")
    args = ap.parse_args()
    import torch
    from digilogic_b200 import SimulatorBuilder, netgen
    from digilogic_b200.batch import Batch
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        peak = 6650.0
    gen = netgen.random_dag(n_gates=args.gates, depth=args.depth, n_inputs=args.inputs, n_outputs=args.outputs, seed=args.seed)
    sim = gen.emit(SimulatorBuilder()).build()
    words = args.lanes // 64
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ov = torch.zeros((len(gen.outputs), words), dtype=torch.int64, device="cuda")
    om = torch.zeros_like(ov)
    first = None
    for shape in args.shapes.split(","):
        f = [int(x) for x in shape.split("x")]
        tile, slots, iters, unroll, ctas = f[:5]
        ld_mode = f[5] if len(f) > 5 else 0
        sweep = f[6] if len(f) > 6 else 1
        hints = f[7] if len(f) > 7 else 0
        discard = f[8] if len(f) > 8 else 1
        pdl = f[9] if len(f) > 9 else 0
        sim.set_option(sim.OPT_PDL, pdl)
        sim.set_option(sim.OPT_L2_HINTS, hints)
        sim.set_option(sim.OPT_VALID_ELISION, args.elision)
        sim.set_option(sim.OPT_EVAL_UNROLL, args.eval_unroll)
        sim.set_option(sim.OPT_PAIR_FUSION, f[10] if len(f) > 10 else 2)
        b = Batch(sim, gen.inputs, gen.outputs, tile, slots, device=0, stream=stream.cuda_stream)
        b.set_option(Batch.OPT_TASK_ITERS, iters)
        b.set_option(Batch.OPT_UNROLL, unroll)
        if ctas:
            b.set_option(Batch.OPT_CTAS_PER_SM, ctas)
        b.set_option(Batch.OPT_SWEEP, sweep)
        b.set_option(Batch.OPT_DISCARD, discard)
        b.run_generated(args.seed, 0, 0, words, 10_000, ov, om)
        b.wait()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            b.run_generated(args.seed, 0, 0, words, 10_000, ov, om)
        e1.record(stream)
        b.wait()
        ms = e0.elapsed_time(e1) / args.steps
        cs = int((ov.sum() ^ om.sum()).item()) & 0xFFFFFFFFFFFFFFFF
        first = cs if first is None else first
        info = b.info
        print(json.dumps({"tile": tile, "slots": slots, "iters": iters, "unroll": unroll, "ctas_per_sm": ctas, "ld_mode": ld_mode, "sweep": sweep, "l2_hints": hints, "discard": discard, "pdl": pdl, "pair_launches": info.pair_launches,
                          "ms": round(ms, 2), "evals_per_s": args.gates * args.lanes / (ms * 1e-3),
                          "algo_frac": 48.0 * args.gates * words / (ms * 1e-3) / 1e9 / peak, "checksum_equal": cs == first,
                          "checksum": f"{cs:#018x}"}), flush=True)
        del b


if __name__ == "__main__":
    main()
