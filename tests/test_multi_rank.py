"""The N>1 path on CPU: world_size-2 gloo processes exercise the lane sharding, the output-plane
all-gather and the result reduction of digilogic_b200.batch (host logic only — the per-rank
planes are stand-ins produced by the CPU oracle, no engine call is made without a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_planes(word_begin, n_words):
    """Output planes of a small DAG for global lane-words [word_begin, word_begin + n_words)."""
    sys.path.insert(0, ROOT)
    import oracle
    from digilogic_b200 import netgen
    gen = netgen.random_dag(n_gates=400, depth=8, n_inputs=16, n_outputs=8, seed=11)
    bs = oracle.BitSliced(gen.emit(oracle.Builder()), 64 * n_words, gen.n_wires)
    v, m = netgen.stimulus_words(11, 16, n_words, word_begin, 1)
    for i, w in enumerate(gen.inputs):
        bs.set_drive(int(w), v[i], m[i])
    assert not bs.eval_dag().any()
    pv = np.stack([bs.get(int(w))[0] for w in gen.outputs])
    pm = np.stack([bs.get(int(w))[1] for w in gen.outputs])
    return pv, pm


def _worker(rank, world, port, total_words, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from digilogic_b200.batch import gather_output_planes, reduce_worst, shard_words
    from digilogic_b200.gsim import SimulationRunResult
    w0, w1 = shard_words(total_words, world, rank)
    pv, pm = _oracle_planes(w0, w1 - w0)
    lv = torch.from_numpy(pv.view(np.int64).copy())
    lm = torch.from_numpy(pm.view(np.int64).copy())
    gv, gm = gather_output_planes(lv, lm, world)
    worst = reduce_worst(SimulationRunResult.MaxStepsReached if rank == 1 else SimulationRunResult.Ok, torch.device("cpu"))
    q.put((rank, gv.numpy().view(np.uint64), gm.numpy().view(np.uint64), int(worst)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2_gloo():
    from digilogic_b200.batch import plan_tiles, shard_words
    assert shard_words(64, 2, 0) == (0, 32) and shard_words(64, 2, 1) == (32, 64)
    with pytest.raises(ValueError):
        shard_words(65, 2, 0)
    assert plan_tiles(10, 4) == [(0, 4), (4, 4), (8, 2)]
    world, total_words = 2, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_words, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_v, ref_m = _oracle_planes(0, total_words)     # the un-sharded batch
    for rank, gv, gm, worst in results:
        assert worst == 1, "MAX over ranks: MaxStepsReached on rank 1 must reach every rank"
        # [world, n_out, words_local] rank-major == lane-word order of the whole batch
        whole_v = np.concatenate([gv[r] for r in range(world)], axis=1)
        whole_m = np.concatenate([gm[r] for r in range(world)], axis=1)
        assert (whole_v == ref_v).all() and (whole_m == ref_m).all()
