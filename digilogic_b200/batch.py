"""Lane-tiled batch runs and stimulus-sharded multi-GPU runs: a ctypes shim over the C ABI's batch runner.

A batch of L stimulus lanes is L independent gsim simulators of one netlist.  Everything that moves data or launches work
— the tiling of the lanes, the slots that work on alternate tiles, the persistent level-sweep kernel, the replicas for
netlists with feedback, the gather of the output planes into every GPU over NVLink or NCCL — lives behind
``b2_batch_*`` / ``b2_comm_*`` (include/digilogic_b200.h, digilogic_b200/csrc/batch.cu).  What is left here is argument
marshalling, the pure arithmetic of sharding and tile choice, and the two small blobs a host has to pass between ranks
(NCCL unique id, CUDA IPC handles), for which this harness uses torch.distributed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .gsim import B200Error, Simulator, SimulationRunResult, _check, _p


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


def auto_batch_shape(sim: Simulator, n_words: int) -> tuple[int, int]:
    """(tile_words, n_slots) for the batch runner: three slots — with three tiles in flight the launch gaps and kernel tails
    of one tile's level sweep are filled by the others — and a tile small enough that TWO levels of all three slots
    (gates per level x 16 B x tile_words x 3 x 2) stay in L2: the next level reads the rows a level wrote from L2, and rows
    that are dead after that are dropped from L2 before they are written back (B2_BATCH_OPT_DISCARD).  Config 3 (5000 gates
    per level) -> 128 x 3 (measured optimum, profiles/README.md).  Where the tile is already at its floor of 32 words and a level
    still exceeds that budget (config 5: 50 000 gates per level = 25.6 MB), two slots do better than three (865 vs 892 ms per
    2^20 lanes of config 5): -> 32 x 2."""
    tile = auto_tile_words(sim, n_words, l2_level_bytes=10 << 20)
    info = sim.info
    per_level = (info.n_gates2 + info.n_gates_wide) / max(info.depth, 1)
    return tile, (2 if per_level * 16 * tile > (16 << 20) else 3)


class _DeviceArray:
    """A device buffer owned by the C library, exposed through __cuda_array_interface__ (torch.as_tensor accepts it)."""

    def __init__(self, ptr: int, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i8", "data": (int(ptr), False), "version": 2}
        self._owner = owner


def _planes(value, valid):
    """(value_ptr, valid_ptr, stride_words) of a pair of [n, words] planes: torch tensors (device or pinned host), numpy
    arrays, or None."""
    if value is None:
        return None, None, 0
    if hasattr(value, "data_ptr"):
        assert value.dim() == 2 and valid.dim() == 2 and value.stride(1) == 1 and valid.stride(0) == value.stride(0)
        return value.data_ptr(), valid.data_ptr(), value.stride(0)
    assert value.ndim == 2 and value.strides[1] == 8 and valid.strides[0] == value.strides[0]
    return value.ctypes.data, valid.ctypes.data, value.strides[0] // 8


class Comm:
    """b2_comm: an NCCL communicator made by the C library (ncclGetUniqueId / ncclCommInitRank)."""

    def __init__(self, n_ranks: int, rank: int, unique_id: bytes, device: int):
        self._lib = _ffi.lib()
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self._lib.b2_comm_init_rank(C.byref(h), n_ranks, rank, buf, device), "b2_comm_init_rank")
        self._h = h
        self.n_ranks, self.rank = n_ranks, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(_ffi.lib().b2_comm_get_unique_id(buf), "b2_comm_get_unique_id")
        return bytes(buf)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b2_comm_free(self._h)
            self._h = None


class Batch:
    """b2_batch: runs lane batches of one compiled netlist (see include/digilogic_b200.h)."""

    OPT_SWEEP, OPT_TASK_ITERS, OPT_UNROLL, OPT_CTAS_PER_SM, OPT_AUTO_GATHER, OPT_DISCARD = 1, 2, 3, 4, 5, 6

    def __init__(self, sim: Simulator, inputs, outputs, tile_words: int, n_slots: int = 1, device: int | None = None,
                 stream: int | None = None):
        self._lib = _ffi.lib()
        self.inputs = np.ascontiguousarray(inputs, np.uint32)
        self.outputs = np.ascontiguousarray(outputs, np.uint32)
        h = C.c_void_p()
        _check(self._lib.b2_batch_new(sim._h, _p(self.inputs), self.inputs.size, _p(self.outputs), self.outputs.size,
                                      int(tile_words), int(n_slots), C.byref(h)), "b2_batch_new")
        self._h = h
        self._comm = None
        if device is not None:
            _check(self._lib.b2_batch_set_device(self._h, device), "b2_batch_set_device")
        if stream is not None:
            _check(self._lib.b2_batch_set_stream(self._h, C.c_void_p(stream)), "b2_batch_set_stream")

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b2_batch_free(self._h)
            self._h = None

    def set_option(self, option: int, value: int):
        _check(self._lib.b2_batch_set_option(self._h, option, value), "b2_batch_set_option")

    @property
    def info(self) -> _ffi.BatchInfo:
        out = _ffi.BatchInfo()
        _check(self._lib.b2_batch_get_info(self._h, C.byref(out)), "b2_batch_get_info")
        return out

    def run_generated(self, seed: int, mode: int, word_begin: int, n_words: int, max_steps: int = 10_000, out_value=None,
                      out_valid=None, out_ring_words: int = 0):
        """Queues one settle of lanes [0, 64 n_words) with device-generated stimulus (global lane-word word_begin + w)."""
        pv, pm, stride = _planes(out_value, out_valid)
        _check(self._lib.b2_batch_run_generated(self._h, seed, mode, word_begin, n_words, max_steps, C.c_void_p(pv), C.c_void_p(pm),
                                                stride, out_ring_words), "b2_batch_run_generated")

    def run_planes(self, in_value, in_valid, n_words: int, max_steps: int = 10_000, out_value=None, out_valid=None,
                   in_ring_words: int = 0, out_ring_words: int = 0):
        """Queues one settle with stimulus planes [n_inputs, words] (device memory, or host memory read in place)."""
        iv, im, istride = _planes(in_value, in_valid)
        pv, pm, stride = _planes(out_value, out_valid)
        _check(self._lib.b2_batch_run_planes(self._h, C.c_void_p(iv), C.c_void_p(im), istride, in_ring_words, n_words, max_steps,
                                             C.c_void_p(pv), C.c_void_p(pm), stride, out_ring_words), "b2_batch_run_planes")

    def wait(self) -> SimulationRunResult:
        worst = C.c_int()
        _check(self._lib.b2_batch_wait(self._h, C.byref(worst)), "b2_batch_wait")
        return SimulationRunResult(worst.value)

    # ---- multi-GPU
    def bind_comm(self, comm: Comm, words_local: int):
        _check(self._lib.b2_batch_bind_comm(self._h, comm._h, words_local), "b2_batch_bind_comm")
        self._comm = comm
        self._words_local = words_local

    def export_gather_handle(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(self._lib.b2_batch_export_gather_handle(self._h, buf), "b2_batch_export_gather_handle")
        return bytes(buf)

    def import_gather_handles(self, handles: list[bytes]):
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(self._lib.b2_batch_import_gather_handles(self._h, buf), "b2_batch_import_gather_handles")

    def gather_outputs(self):
        _check(self._lib.b2_batch_gather_outputs(self._h), "b2_batch_gather_outputs")

    def gathered(self):
        """(value, valid) device arrays [n_ranks, n_outputs, words_local] (int64 storage) of this rank's gathered planes."""
        pv, pm = C.c_void_p(), C.c_void_p()
        _check(self._lib.b2_batch_gathered(self._h, C.byref(pv), C.byref(pm)), "b2_batch_gathered")
        shape = (self._comm.n_ranks, self.outputs.size, self._words_local)
        return _DeviceArray(pv.value, shape, self), _DeviceArray(pm.value, shape, self)


def setup_comm(batch: Batch, words_local: int, device: int, map_peers: bool = True) -> Comm:
    """Creates the C library's communicator over the ranks of an initialised torch.distributed group and binds it: the
    NCCL unique id travels rank 0 -> all, the CUDA IPC handles of the gathered planes all -> all (object collectives)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = Comm(world, rank, box[0], device)
    batch.bind_comm(comm, words_local)
    if map_peers and world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, batch.export_gather_handle())
        batch.import_gather_handles(handles)
    return comm


# ------------------------------------------------------------------------------------------------------------------
# Harness helpers kept from round 1 (same call shapes; the work happens in the C library)

class BatchRunner:
    """One slot, tile by tile: the round-1 call shape over b2_batch."""

    def __init__(self, sim: Simulator, inputs, outputs, tile_words: int, max_steps: int = 10_000, n_slots: int = 1,
                 stream: int | None = None):
        self.sim = sim
        self.tile_words = int(tile_words)
        self.max_steps = int(max_steps)
        self.batch = Batch(sim, inputs, outputs, tile_words, n_slots, stream=stream)
        self.inputs, self.outputs = self.batch.inputs, self.batch.outputs

    def run_generated(self, word_begin: int, n_words: int, seed: int, mode: int, out_value, out_valid, out_word0: int = 0):
        ov = out_value[:, out_word0:] if out_word0 else out_value
        om = out_valid[:, out_word0:] if out_word0 else out_valid
        self.batch.run_generated(seed, mode, word_begin, n_words, self.max_steps, ov, om)
        return self.batch.wait()

    def run_host(self, in_value, in_valid, out_value, out_valid, n_words: int | None = None):
        ring = in_value.shape[1]
        n_words = ring if n_words is None else n_words
        self.batch.run_planes(in_value, in_valid, n_words, self.max_steps, out_value, out_valid,
                              in_ring_words=ring if n_words > ring else 0, out_ring_words=ring if n_words > ring else 0)
        return self.batch.wait()


def gather_output_planes(local_value, local_valid, world_size: int):
    """torch.distributed all-gather of packed output planes [n_out, words_local] -> [world, n_out, words_local] (the
    rank-major layout of the C library's gathered planes).  Used by the CPU (gloo) test of the sharding arithmetic and as
    the independent check of b2_batch_gather_outputs on GPUs; the product path gathers inside the C library."""
    import torch
    import torch.distributed as dist

    if world_size == 1:
        return local_value.unsqueeze(0), local_valid.unsqueeze(0)
    gv = torch.empty((world_size,) + tuple(local_value.shape), dtype=local_value.dtype, device=local_value.device)
    gm = torch.empty_like(gv)
    dist.all_gather_into_tensor(gv.view(-1, local_value.shape[-1]), local_value.contiguous())
    dist.all_gather_into_tensor(gm.view(-1, local_valid.shape[-1]), local_valid.contiguous())
    return gv, gm


def reduce_worst(worst: SimulationRunResult, device) -> SimulationRunResult:
    """MAX all-reduce of the run result (conflict > max-steps > ok) over ranks (torch.distributed; harness only)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return worst
    t = torch.tensor([int(worst)], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return SimulationRunResult(int(t.item()))


__all__ = ["Batch", "BatchRunner", "Comm", "B200Error", "auto_batch_shape", "auto_tile_words", "choose_tile_words",
           "gather_output_planes", "plan_tiles", "reduce_worst", "setup_comm", "shard_words"]
