"""The batch runner behind the C ABI (b2_batch_*): the persistent level-sweep kernel and the replica path, against the CPU
oracle on the same seeded inputs — every designated wire, every lane — and against each other."""
import os
import subprocess
import sys

import numpy as np
import pytest

import digilogic_b200 as d
import oracle
from digilogic_b200 import netgen
from digilogic_b200.batch import Batch, BatchRunner, Comm
from digilogic_b200.gsim import SimulationRunResult

from helpers import ALL1, random_netlist, stimulus_words

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_dag(nl_or_gen, n_wires, inputs, words, seed, word_begin, mode):
    ob = nl_or_gen.replay(oracle.Builder()) if hasattr(nl_or_gen, "replay") else nl_or_gen.emit(oracle.Builder())
    bs = oracle.BitSliced(ob, 64 * words, n_wires)
    v, m = stimulus_words(seed, len(inputs), words, word_begin, mode)
    for i, w in enumerate(inputs):
        bs.set_drive(int(w), v[i], m[i])
    assert not bs.eval_dag().any()
    return bs.get_all()


@pytest.mark.parametrize("tile,slots,iters,unroll", [(2, 1, 1, 4), (8, 3, 1, 2), (32, 4, 2, 4), (64, 5, 1, 4), (128, 2, 3, 2), (512, 2, 1, 4)])
def test_sweep_matches_oracle_on_every_wire(tile, slots, iters, unroll):
    """Mixed classes (one/two/three-input gates, n-ary gates, multiplexers), 4-state stimulus, a batch that is neither a
    whole number of tiles nor of slot groups: every wire of every lane equals the dense CPU model."""
    rng = np.random.default_rng(tile * 131 + slots)
    nl, first_in = random_netlist(rng, 24, 700, cyclic=False, wide=True, mux=True)
    inputs = np.arange(first_in, first_in + 24, dtype=np.uint32)
    all_wires = np.arange(nl.n_wires, dtype=np.uint32)
    sim = nl.replay(d.SimulatorBuilder()).build()
    assert not sim.info.cyclic
    words = tile * (2 * slots + 1) + (tile // 2 + 1 if tile > 2 else 1)      # short last tile, partial last group
    b = Batch(sim, inputs, all_wires, tile, slots)
    b.set_option(Batch.OPT_SWEEP, 1)
    b.set_option(Batch.OPT_TASK_ITERS, iters)
    b.set_option(Batch.OPT_UNROLL, unroll)
    ov = np.zeros((nl.n_wires, words), np.uint64)
    om = np.zeros_like(ov)
    b.run_generated(0xABCD, 1, 77, words, 10_000, ov, om)
    assert b.wait() == SimulationRunResult.Ok
    assert b.info.sweep == 1 and b.info.last_tiles == -(-words // b.info.tile_words)
    rv, rm = _oracle_dag(nl, nl.n_wires, inputs, words, 0xABCD, 77, 1)
    assert (ov == rv).all() and (om == rm).all()
    # a second run on the same slots (their rows are reused) with other stimulus
    b.run_generated(0x1234, 0, 5, words, 10_000, ov, om)
    assert b.wait() == SimulationRunResult.Ok
    rv, rm = _oracle_dag(nl, nl.n_wires, inputs, words, 0x1234, 5, 0)
    assert (ov == rv).all() and (om == rm).all()


def test_sweep_equals_levelised_engine_and_replicas():
    """100k-gate DAG: the sweep, the replica path of the batch runner and b2_run_sim_lanes give the same output planes
    (tiling invariance: lanes are independent simulators, the stimulus is a function of the global lane-word)."""
    import torch
    gen = netgen.random_dag(n_gates=100_000, depth=50, n_inputs=512, n_outputs=256)
    sim = gen.emit(d.SimulatorBuilder()).build()
    words = 1024
    outs = []
    for tile, slots, sweep in ((64, 4, 1), (256, 2, 1), (128, 3, 0)):
        b = Batch(sim, gen.inputs, gen.outputs, tile, slots)
        b.set_option(Batch.OPT_SWEEP, sweep)
        ov = torch.zeros((len(gen.outputs), words), dtype=torch.int64, device="cuda")
        om = torch.zeros_like(ov)
        b.run_generated(0xD1610003, 1, 1000, words, 10_000, ov, om)
        assert b.wait() == SimulationRunResult.Ok
        assert b.info.sweep == sweep
        outs.append((ov.cpu(), om.cpu()))
    one = sim.clone()
    one.set_lanes(64 * words)
    one.fill_stimulus(gen.inputs, 0xD1610003, lane_word_base=1000, mode=1)
    res, _, _ = one.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok and one.info.last_mode == 1
    gv, gm = one.gather_wires_host(gen.outputs)
    for ov, om in outs:
        assert (ov.numpy().view(np.uint64) == gv).all() and (om.numpy().view(np.uint64) == gm).all()
    assert (gm != ALL1).any(), "4-state stimulus must leave some X outputs"


@pytest.mark.parametrize("sweep", [1, 0])
def test_planes_from_device_pinned_and_pageable_host(sweep):
    import torch
    gen = netgen.random_dag(n_gates=20_000, depth=20, n_inputs=128, n_outputs=64, seed=5)
    sim = gen.emit(d.SimulatorBuilder()).build()
    words, tile = 96, 32
    v, m = stimulus_words(99, len(gen.inputs), words, 0, 1)
    rv, rm = _oracle_dag(gen, gen.n_wires, gen.inputs, words, 99, 0, 1)
    want_v, want_m = rv[gen.outputs], rm[gen.outputs]
    b = Batch(sim, gen.inputs, gen.outputs, tile, 3)
    b.set_option(Batch.OPT_SWEEP, sweep)
    # pageable host in, pageable host out (the sweep page-locks them for the run; the replicas stage them)
    ov, om = np.zeros((64, words), np.uint64), np.zeros((64, words), np.uint64)
    b.run_planes(v, m, words, 10_000, ov, om)
    assert b.wait() == SimulationRunResult.Ok
    assert (ov == want_v).all() and (om == want_m).all()
    # pinned host in, device out
    pv = torch.from_numpy(v.view(np.int64)).pin_memory()
    pm = torch.from_numpy(m.view(np.int64)).pin_memory()
    dv = torch.zeros((64, words), dtype=torch.int64, device="cuda")
    dm = torch.zeros_like(dv)
    b.run_planes(pv, pm, words, 10_000, dv, dm)
    assert b.wait() == SimulationRunResult.Ok
    assert (dv.cpu().numpy().view(np.uint64) == want_v).all() and (dm.cpu().numpy().view(np.uint64) == want_m).all()
    assert b.info.sweep == sweep
    # device in, pinned host out; odd strides (scalar path)
    gv = torch.zeros((128, words + 1), dtype=torch.int64, device="cuda")
    gm = torch.zeros_like(gv)
    gv[:, :words] = torch.from_numpy(v.view(np.int64)).cuda()
    gm[:, :words] = torch.from_numpy(m.view(np.int64)).cuda()
    hv = torch.zeros((64, words + 3), dtype=torch.int64).pin_memory()
    hm = torch.zeros((64, words + 3), dtype=torch.int64).pin_memory()
    b.run_planes(gv, gm, words, 10_000, hv, hm)
    assert b.wait() == SimulationRunResult.Ok
    assert (hv.numpy().view(np.uint64)[:, :words] == want_v).all() and (hm.numpy().view(np.uint64)[:, :words] == want_m).all()
    # rings: the stimulus planes hold one tile group and are re-read, the output ring keeps the last lap
    ring = 2 * tile
    rv_, rm_ = v[:, :ring].copy(), m[:, :ring].copy()      # (kept alive: the run reads them after the call has returned)
    b.run_planes(rv_, rm_, 3 * ring, 10_000, ov[:, :ring], om[:, :ring], in_ring_words=ring, out_ring_words=ring)
    assert b.wait() == SimulationRunResult.Ok
    assert (ov[:, :ring] == want_v[:, :ring]).all() and (om[:, :ring] == want_m[:, :ring]).all()


def test_replica_path_runs_feedback_netlists_with_per_lane_results():
    """A cyclic netlist cannot be swept: the batch runner walks it tile by tile on replicas with the unit-delay semantics
    of b2_run_sim_lanes; outputs and the worst result equal the scalar oracle's."""
    rng = np.random.default_rng(7)
    nl, first_in = random_netlist(rng, 12, 120, cyclic=True)
    inputs = np.arange(first_in, first_in + 12, dtype=np.uint32)
    all_wires = np.arange(nl.n_wires, dtype=np.uint32)
    sim = nl.replay(d.SimulatorBuilder()).build()
    assert sim.info.cyclic
    tile, words, max_steps = 2, 6, 40
    b = Batch(sim, inputs, all_wires, tile, 2)
    assert b.info.sweep == 0
    ov, om = np.zeros((nl.n_wires, words), np.uint64), np.zeros((nl.n_wires, words), np.uint64)
    b.run_generated(31, 1, 0, words, max_steps, ov, om)
    worst = b.wait()
    lanes = oracle.Lanes(nl.replay(oracle.Builder()), 64 * words)
    v, m = stimulus_words(31, 12, words, 0, 1)
    for i, w in enumerate(inputs):
        lanes.set_drive(int(w), v[i], m[i])
    status, _, _ = lanes.run(max_steps)
    assert int(worst) == int(status.max())
    rv, rm = lanes.get_all(nl.n_wires)
    ok = status != 2                                   # the state after a conflict is outside the contract
    okw = np.packbits(ok.reshape(words, 64), axis=1, bitorder="little").view(np.uint64).reshape(words)
    assert ((ov ^ rv) & okw == 0).all() and ((om ^ rm) & okw == 0).all()


def test_clone_shares_the_netlist_and_runs_independently():
    gen = netgen.random_dag(n_gates=4000, depth=10, n_inputs=32, n_outputs=16, seed=3)
    a = gen.emit(d.SimulatorBuilder()).build()
    a.set_lanes(128)
    c = a.clone()
    c.set_lanes(64)
    a.fill_stimulus(gen.inputs, 1, 0, 0)
    c.fill_stimulus(gen.inputs, 1, 1, 0)       # the clone simulates lane-word 1
    assert a.run_sim_lanes(100)[0] == SimulationRunResult.Ok and c.run_sim_lanes(100)[0] == SimulationRunResult.Ok
    av, am = a.gather_wires_host(gen.outputs)
    cv, cm = c.gather_wires_host(gen.outputs)
    assert (av[:, 1] == cv[:, 0]).all() and (am[:, 1] == cm[:, 0]).all()
    assert a.info.n_rows == c.info.n_rows


def test_single_rank_comm_gathers_into_own_planes():
    """b2_comm_* with one rank (NCCL is loaded, a communicator is made) and the gathered planes of b2_batch_gathered."""
    import torch
    gen = netgen.random_dag(n_gates=4000, depth=10, n_inputs=32, n_outputs=16, seed=3)
    sim = gen.emit(d.SimulatorBuilder()).build()
    words = 64
    b = Batch(sim, gen.inputs, gen.outputs, 32, 2, device=0)
    b.set_option(Batch.OPT_SWEEP, 1)
    comm = Comm(1, 0, Comm.unique_id(), 0)
    b.bind_comm(comm, words)
    b.run_generated(9, 0, 0, words)          # with a communicator bound, a run ends with b2_batch_gather_outputs
    assert b.wait() == SimulationRunResult.Ok
    gv, gm = (torch.as_tensor(x, device="cuda") for x in b.gathered())
    rv, rm = _oracle_dag(gen, gen.n_wires, gen.inputs, words, 9, 0, 0)
    assert (gv[0].cpu().numpy().view(np.uint64) == rv[gen.outputs]).all()
    assert (gm[0].cpu().numpy().view(np.uint64) == rm[gen.outputs]).all()


def test_batch_argument_errors():
    gen = netgen.random_dag(n_gates=400, depth=4, n_inputs=8, n_outputs=4, seed=3)
    sim = gen.emit(d.SimulatorBuilder()).build()
    with pytest.raises(d.gsim.ComponentError):
        Batch(sim, [10_000], gen.outputs, 32, 2)           # InvalidWireId
    with pytest.raises(d.gsim.B200Error):
        Batch(sim, gen.inputs, gen.outputs, 0, 2)
    b = Batch(sim, gen.inputs, gen.outputs, 32, 2)
    ov = np.zeros((4, 16), np.uint64)
    with pytest.raises(d.gsim.B200Error):
        b.run_generated(1, 0, 0, 64, 100, ov, ov.copy())   # planes shorter than the batch
    with pytest.raises(d.gsim.B200Error):
        b.run_generated(1, 0, 0, 64, 100, ov, ov.copy(), out_ring_words=24)   # a ring that tiles would straddle


def _two_rank_worker(rank, world, port, words_local, map_peers, sweep, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from digilogic_b200.batch import Batch, gather_output_planes, setup_comm, shard_words
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    gen = netgen.random_dag(n_gates=20_000, depth=20, n_inputs=128, n_outputs=64, seed=5)
    sim = gen.emit(d.SimulatorBuilder()).build()
    b = Batch(sim, gen.inputs, gen.outputs, 32, 3, device=rank)
    b.set_option(Batch.OPT_SWEEP, sweep)
    setup_comm(b, words_local, rank, map_peers=map_peers)
    w0, _ = shard_words(words_local * world, world, rank)
    lv = torch.zeros((64, words_local), dtype=torch.int64, device="cuda")
    lm = torch.zeros_like(lv)
    for seed in (21, 22):                                   # two steps: the second overwrites planes peers have read
        b.run_generated(seed, 1, w0, words_local, 10_000, lv, lm)
        assert b.wait() == 0
        gv, gm = (torch.as_tensor(x, device="cuda") for x in b.gathered())
        tv, tm = gather_output_planes(lv, lm, world)        # torch.distributed's all-gather of the same local planes
        assert torch.equal(gv, tv) and torch.equal(gm, tm)
    q.put((rank, b.info.peers_mapped, gv.cpu().numpy().view(np.uint64), gm.cpu().numpy().view(np.uint64)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("map_peers,sweep", [(True, 0), (False, 0), (True, 1), (False, 1)])
def test_two_gpus_gathered_planes_equal_the_unsharded_batch(map_peers, sweep):
    """N = 2 on hardware: lane-sharded runs, output rows stored into the peer's planes over NVLink by the sweep kernel
    (map_peers) or moved by ncclAllGather, on either engine; both equal torch.distributed's all-gather and the unsharded oracle batch."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world, words_local = 2, 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, world, port, words_local, map_peers, sweep, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gen = netgen.random_dag(n_gates=20_000, depth=20, n_inputs=128, n_outputs=64, seed=5)
    rv, rm = _oracle_dag(gen, gen.n_wires, gen.inputs, words_local * world, 22, 0, 1)
    for rank, mapped, gv, gm in results:
        assert mapped == (1 if map_peers else 0)
        whole_v = np.concatenate([gv[r] for r in range(world)], axis=1)
        whole_m = np.concatenate([gm[r] for r in range(world)], axis=1)
        assert (whole_v == rv[gen.outputs]).all() and (whole_m == rm[gen.outputs]).all()


def test_plain_c_client_drives_a_two_slot_batch(tmp_path):
    """tests/c_batch_smoke.c: a C program builds a netlist, runs a 2-slot batch from host planes and checks the outputs
    against its own evaluation — the batch path needs nothing but the C ABI."""
    exe = str(tmp_path / "c_batch_smoke")
    lib_dir = os.path.join(ROOT, "digilogic_b200", "lib")
    res = subprocess.run(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_batch_smoke.c"),
                          "-L", lib_dir, "-ldigilogic_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "c_batch_smoke ok" in run.stdout, run.stdout + run.stderr
