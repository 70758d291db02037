#!/usr/bin/env python
"""Experiment: K simulator replicas on K CUDA streams working on alternate lane tiles of config 3,
so that launch gaps and kernel tails of one tile are filled by another and small tiles (whose
level outputs stay in L2) become practical.  Prints one JSON line per (tile, K)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from digilogic_b200 import SimulatorBuilder, netgen  # noqa: E402
from digilogic_b200.batch import BatchRunner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", default="512,1024")
ap.add_argument("--ks", default="1,2,3")
ap.add_argument("--lanes", type=int, default=1 << 24)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--unroll", default="0")
ap.add_argument("--elision", type=int, default=0)
ap.add_argument("--threads", type=int, default=1)
args = ap.parse_args()

gen = netgen.random_dag()
dev = torch.device("cuda", 0)
words = args.lanes // 64
main = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(main)
out_v = torch.zeros((len(gen.outputs), words), dtype=torch.int64, device=dev)
out_m = torch.zeros_like(out_v)
ref = None
for tile in [int(x) for x in args.tiles.split(",")]:
  for unroll in [int(x) for x in args.unroll.split(",")]:
    for k in [int(x) for x in args.ks.split(",")]:
        from digilogic_b200.batch import InterleavedRunner

        def make_sim():
            sim = gen.emit(SimulatorBuilder()).build()
            sim.set_option(sim.OPT_LEVEL_GRAPH, args.graph)
            sim.set_option(sim.OPT_EVAL_UNROLL, unroll)
            sim.set_option(sim.OPT_VALID_ELISION, args.elision)
            return sim

        runners = InterleavedRunner(make_sim, gen.inputs, gen.outputs, tile, k, device=dev, host_threads=bool(args.threads))

        def step():
            runners.run_generated(0, words, 0xD1610003, 0, out_v, out_m, 0)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _ in range(args.steps):
            step()
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        cs = int((out_v.sum() ^ out_m.sum()).item()) & 0xFFFFFFFFFFFFFFFF
        ref = ref or cs
        evals = 1e6 * args.lanes / (ms * 1e-3)
        print(json.dumps({"tile_words": tile, "streams": k, "ms_per_step": ms, "gate_lane_evals_per_s": evals,
                          "graph": args.graph, "threads": args.threads, "elision": args.elision, "unroll": unroll, "frac_of_measured_hbm": 0.75 * evals / 6554.9e9, "checksum_ok": cs == ref}), flush=True)
        del runners
        torch.cuda.empty_cache()
