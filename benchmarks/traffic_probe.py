#!/usr/bin/env python
"""One settle of BASELINE config 3 inside a cudaProfilerStart/Stop range, for `ncu --replay-mode range|app-range`: DRAM bytes
and L2 hit rate of a whole step WITH all replicas running concurrently (per-kernel ncu captures serialise the streams and
so see one replica's working set in L2).  Plain run prints the step time; under ncu the range's metrics are in the report.

    ncu --replay-mode app-range --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
        --csv --log-file gpurun_out/traffic.csv python benchmarks/traffic_probe.py --lanes 2097152
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
    ap.add_argument("--lanes", type=int, default=1 << 21)
    ap.add_argument("--tile-words", type=int, default=256)
    ap.add_argument("--slots", type=int, default=3)
    ap.add_argument("--sweep", type=int, default=0)
    ap.add_argument("--pair", type=int, default=2)
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0xD1610003)
    args = ap.parse_args()
    import torch
    from digilogic_b200 import SimulatorBuilder, netgen
    from digilogic_b200.batch import Batch
    gen = netgen.random_dag(n_gates=args.gates, depth=args.depth, n_inputs=args.inputs, n_outputs=args.outputs, seed=args.seed)
    sim = gen.emit(SimulatorBuilder()).build()
    words = args.lanes // 64
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ov = torch.zeros((len(gen.outputs), words), dtype=torch.int64, device="cuda")
    om = torch.zeros_like(ov)
    sim.set_option(sim.OPT_PAIR_FUSION, args.pair)
    b = Batch(sim, gen.inputs, gen.outputs, args.tile_words, args.slots, device=0, stream=stream.cuda_stream)
    b.set_option(Batch.OPT_SWEEP, args.sweep)
    for _ in range(2):                       # warm-up: graphs captured, stores allocated
        b.run_generated(args.seed, 0, 0, words, 10_000, ov, om)
    b.wait()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record(stream)
    b.run_generated(args.seed, 0, 0, words, 10_000, ov, om)
    e1.record(stream)
    b.wait()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    gate_words = args.gates * words
    print(json.dumps({"lanes": args.lanes, "tile_words": args.tile_words, "slots": args.slots, "sweep": args.sweep, "pair_launches": b.info.pair_launches, "ms": ms,
                      "gate_words": gate_words, "algorithmic_bytes": 48 * gate_words,
                      "checksum": f"{int((ov.sum() ^ om.sum()).item()) & 0xFFFFFFFFFFFFFFFF:#018x}"}))


if __name__ == "__main__":
    main()
