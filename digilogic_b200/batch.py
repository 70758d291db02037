"""Lane-tiled batch runs and stimulus-sharded multi-GPU runs.

A batch of L stimulus lanes is L independent gsim simulators of one netlist.  One GPU holds a
full netlist replica and a wire-state store of `tile_words` 64-lane words; the batch is
processed tile by tile (stimulus in -> settle -> designated output rows out).  Across GPUs the
lane-words are split into contiguous ranges, one per rank, with no traffic between GPUs until
the packed output planes are all-gathered once at the end (NCCL over NVLink; `gloo` in the CPU
tests).  torch is used for device buffers, streams and the collective only.
"""
from __future__ import annotations

import numpy as np

from .gsim import Simulator, SimulationRunResult


def shard_words(total_words: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous lane-word range [begin, end) owned by `rank` (equal ranges; total must divide)."""
    if total_words % world_size:
        raise ValueError(f"{total_words} lane-words do not split evenly over {world_size} ranks")
    per = total_words // world_size
    return rank * per, (rank + 1) * per


def plan_tiles(n_words: int, tile_words: int) -> list[tuple[int, int]]:
    """[(first_word, n_words)] tiles covering n_words; the last tile may be short."""
    return [(w, min(tile_words, n_words - w)) for w in range(0, n_words, tile_words)]


def choose_tile_words(sim: Simulator, n_words: int, budget_bytes: int, unit_delay: bool = False) -> int:
    """Largest tile (multiple of 512 words where possible) whose store fits the byte budget."""
    info = sim.info
    per_word = (info.n_rows * (2 if unit_delay else 1) + info.n_sources) * 16 + 64
    t = max(int(budget_bytes // per_word), 1)
    if t >= 512:
        t -= t % 512
    return min(t, n_words)


def auto_tile_words(sim: Simulator, n_words: int, l2_level_bytes: int = 20 << 20) -> int:
    """Tile size (64-lane words) for the levelised path: small enough that one level's output rows
    (gates per level x tile x 16 B) stay in L2 until the next level has read them — about 20 MB per
    replica with two replicas in flight on a 126 MB L2 — but never below 32 words (one warp per row
    pair keeps loads coalesced) nor above what the batch holds.  Config 3 (5000 gates per level)
    -> 256, config 5 (50 000 per level) -> 32; measured optimum in profiles/README.md."""
    info = sim.info
    if info.cyclic or info.depth == 0:
        return min(n_words, 1024)
    per_level = max((info.n_gates2 + info.n_gates_wide) / info.depth, 1.0)
    t = int(l2_level_bytes / (16.0 * per_level))
    t = max(32, min(t, 4096))
    t = 1 << (t.bit_length() - 1)          # power of two: divides the usual batch sizes
    while t > 1 and n_words % t:
        t >>= 1
    return max(1, min(t, n_words))


class BatchRunner:
    """Runs a lane batch through one simulator replica, tile by tile."""

    def __init__(self, sim: Simulator, inputs, outputs, tile_words: int, max_steps: int = 10_000):
        self.sim = sim
        self.inputs = np.ascontiguousarray(inputs, np.uint32)
        self.outputs = np.ascontiguousarray(outputs, np.uint32)
        self.tile_words = int(tile_words)
        self.max_steps = int(max_steps)
        sim.set_lanes(self.tile_words * 64)

    # -- device-resident stimulus (generated on the GPU from the global lane-word index)
    def run_generated(self, word_begin: int, n_words: int, seed: int, mode: int, out_value, out_valid,
                      out_word0: int = 0, record_events=None) -> SimulationRunResult:
        """out_value/out_valid: device tensors [n_outputs, >= out_word0 + n_words] (int64 storage)."""
        assert n_words % self.tile_words == 0, "batch must be a whole number of tiles"
        worst = SimulationRunResult.Ok
        stride = out_value.stride(0)
        base_v, base_m = out_value.data_ptr(), out_valid.data_ptr()
        for w0, _ in plan_tiles(n_words, self.tile_words):
            self.sim.fill_stimulus(self.inputs, seed, lane_word_base=word_begin + w0, mode=mode)
            if record_events is not None:
                record_events("start")
            r = self.sim.run_sim_lanes_async(self.max_steps)
            if record_events is not None:
                record_events("end")
            worst = max(worst, r)
            off = (out_word0 + w0) * 8
            self.sim.gather_wires_ptr(self.outputs, base_v + off, base_m + off, stride, device=True)
        return worst

    # -- host stimulus through the C ABI (end-to-end path: H2D copies and D2H read-back included)
    def run_host(self, in_value, in_valid, out_value, out_valid, n_words: int | None = None,
                 dev_out_value=None, dev_out_valid=None) -> SimulationRunResult:
        """in_*: host arrays [n_inputs, ring_words]; out_*: host arrays [n_outputs, ring_words] (numpy
        views of pinned memory).  n_words lane-words are processed; if that is more than the host
        arrays hold, the arrays are used as a ring (tile i reads/writes slot i mod ring).  The drives
        of tile i+1 are uploaded while tile i settles, and the outputs of tile i are copied out
        while tile i+1 settles (the engine's copy streams); the call returns after the last copy.
        dev_out_*: optional device tensors [n_outputs, n_words] that also receive the output planes
        (multi-GPU runs all-gather those with NCCL)."""
        ring_words = in_value.shape[1]
        n_words = ring_words if n_words is None else n_words
        assert n_words % self.tile_words == 0 and ring_words % self.tile_words == 0
        assert out_value.shape[1] == ring_words
        worst = SimulationRunResult.Ok
        in_stride, out_stride = in_value.strides[0] // 8, out_value.strides[0] // 8
        tiles = plan_tiles(n_words, self.tile_words)

        def upload(w0):
            o = 8 * (w0 % ring_words)
            self.sim.set_drives_lanes_ptr(self.inputs, in_value.ctypes.data + o, in_valid.ctypes.data + o, in_stride)

        upload(tiles[0][0])
        for i, (w0, _) in enumerate(tiles):
            r = self.sim.run_sim_lanes_async(self.max_steps)
            worst = max(worst, r)
            if i + 1 < len(tiles):
                upload(tiles[i + 1][0])
            o = 8 * (w0 % ring_words)
            self.sim.gather_wires_ptr(self.outputs, out_value.ctypes.data + o, out_valid.ctypes.data + o, out_stride,
                                      device=False, async_=True)
            if dev_out_value is not None:
                self.sim.gather_wires_ptr(self.outputs, dev_out_value.data_ptr() + 8 * w0, dev_out_valid.data_ptr() + 8 * w0,
                                          dev_out_value.stride(0), device=True)
        self.sim.synchronize()
        return worst


class InterleavedRunner:
    """K simulator replicas of one netlist on K CUDA streams, taking alternate lane tiles of a batch.

    Small tiles keep a level's output rows in the 126 MB L2 until the next level reads them (a
    third of the fan-in traffic never reaches HBM), but a single stream of small kernels is
    launch-gap bound; with K tiles in flight the gaps and kernel tails of one tile are filled by
    the others.  The replicas share nothing but the GPU: lanes are independent simulators."""

    def __init__(self, make_sim, inputs, outputs, tile_words: int, n_streams: int, max_steps: int = 10_000, device=None,
                 host_threads: bool = True):
        import torch
        from concurrent.futures import ThreadPoolExecutor
        self.torch = torch
        self.pool = ThreadPoolExecutor(max_workers=n_streams) if (host_threads and n_streams > 1) else None
        self.streams = [torch.cuda.Stream(device=device) for _ in range(n_streams)]
        self.runners = []
        for st in self.streams:
            sim = make_sim()
            sim.set_stream(st.cuda_stream)
            self.runners.append(BatchRunner(sim, inputs, outputs, tile_words, max_steps))
        self.tile_words = int(tile_words)

    @property
    def sims(self):
        return [r.sim for r in self.runners]

    def run_generated(self, word_begin: int, n_words: int, seed: int, mode: int, out_value, out_valid,
                      out_word0: int = 0) -> SimulationRunResult:
        """Same contract as BatchRunner.run_generated; work is forked from and joined to the
        caller's current stream, so events recorded there bracket it."""
        assert n_words % self.tile_words == 0
        main = self.torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(main)
        k = len(self.runners)
        tiles = plan_tiles(n_words, self.tile_words)

        def drive(j):
            # host_threads=True (default): one host thread per replica, so that the replicas' launches are issued
            # in parallel (ctypes releases the GIL during C-ABI calls).  A/B on one box, tile 256 x 2 replicas:
            # 1401 ms per step with threads, 1680 ms with one thread issuing both replicas' graphs.
            w = SimulationRunResult.Ok
            for w0, nw in tiles[j::k]:
                w = max(w, self.runners[j].run_generated(word_begin + w0, nw, seed, mode, out_value, out_valid, out_word0 + w0))
            return w

        if self.pool is not None and k > 1:
            worst = max(self.pool.map(drive, range(k)))
        else:
            worst = max(drive(j) for j in range(k))
        for st in self.streams:
            main.wait_stream(st)
        return worst


    def run_host(self, in_value, in_valid, out_value, out_valid, n_words: int | None = None,
                 dev_out_value=None, dev_out_valid=None) -> SimulationRunResult:
        """Same contract as BatchRunner.run_host (host planes in, host planes out, optional ring),
        with the K replicas' upload -> settle -> read-back pipelines interleaved tile by tile."""
        ring_words = in_value.shape[1]
        n_words = ring_words if n_words is None else n_words
        tw = self.tile_words
        assert n_words % tw == 0 and ring_words % tw == 0 and out_value.shape[1] == ring_words
        main = self.torch.cuda.current_stream()
        for st in self.streams:
            st.wait_stream(main)
        in_stride, out_stride = in_value.strides[0] // 8, out_value.strides[0] // 8
        tiles = plan_tiles(n_words, tw)
        k = len(self.runners)

        def drive(j):
            # replica j's pipeline over its tiles: upload(next) while the current tile settles, read-back async
            r = self.runners[j]
            mine = tiles[j::k]
            w = SimulationRunResult.Ok

            def upload(w0):
                o = 8 * (w0 % ring_words)
                r.sim.set_drives_lanes_ptr(r.inputs, in_value.ctypes.data + o, in_valid.ctypes.data + o, in_stride)

            if mine:
                upload(mine[0][0])
            for i, (w0, _) in enumerate(mine):
                w = max(w, r.sim.run_sim_lanes_async(r.max_steps))
                if i + 1 < len(mine):
                    upload(mine[i + 1][0])    # waits for this replica's drives to be consumed, not for its settle
                o = 8 * (w0 % ring_words)
                r.sim.gather_wires_ptr(r.outputs, out_value.ctypes.data + o, out_valid.ctypes.data + o, out_stride,
                                       device=False, async_=True)
                if dev_out_value is not None:
                    r.sim.gather_wires_ptr(r.outputs, dev_out_value.data_ptr() + 8 * w0, dev_out_valid.data_ptr() + 8 * w0,
                                           dev_out_value.stride(0), device=True)
            r.sim.synchronize()
            return w

        if self.pool is not None and k > 1:
            worst = max(self.pool.map(drive, range(k)))
        else:
            worst = max(drive(j) for j in range(k))
        for st in self.streams:
            main.wait_stream(st)
        return worst


def gather_output_planes(local_value, local_valid, world_size: int):
    """All-gather of the packed output planes: [n_out, words_local] per rank ->
    [world_size, n_out, words_local] (rank-major = lane-word order, since ranks own contiguous
    lane-word ranges).  NCCL on GPU tensors, gloo on CPU tensors."""
    import torch
    import torch.distributed as dist

    if world_size == 1:
        return local_value.unsqueeze(0), local_valid.unsqueeze(0)
    gv = torch.empty((world_size,) + tuple(local_value.shape), dtype=local_value.dtype, device=local_value.device)
    gm = torch.empty_like(gv)
    # the output is the concatenation of the ranks' planes along dim 0 (the form gloo also accepts)
    dist.all_gather_into_tensor(gv.view(-1, local_value.shape[-1]), local_value.contiguous())
    dist.all_gather_into_tensor(gm.view(-1, local_valid.shape[-1]), local_valid.contiguous())
    return gv, gm


def reduce_worst(worst: SimulationRunResult, device) -> SimulationRunResult:
    """MAX all-reduce of the run result (conflict > max-steps > ok) over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return worst
    t = torch.tensor([int(worst)], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return SimulationRunResult(int(t.item()))
