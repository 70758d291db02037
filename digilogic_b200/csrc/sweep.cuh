// sweep.cuh — the persistent level-sweep kernel of the batch runner (batch.cu).
//
// One launch settles a whole batch of lane tiles of an acyclic, single-driver netlist (the levelised rule of
// Engine::run_levelised: the fixed point is unique, so it is computed once, level by level): stimulus in -> every level
// -> designated output rows out, for every tile, with no kernel boundary and no host involvement in between.
//
// Work is cut into TASKS: a chunk of the inputs (stimulus), of one level's gates, or of the outputs, for ONE tile.  K
// tiles are in flight at once, each in its own SLOT (a private wire-state store of `T` lane-words per row); tile t runs
// in slot t % K as that slot's pass t / K.  Tasks are numbered ("tickets") pass-group by pass-group, inside a group level
// by level, inside a level slot by slot:
//
//     ticket = ((group * tasks_per_pass) + task0[seg]) * K  +  slot * n_tasks[seg]  +  index
//
// and handed out by one atomic counter, so CTAs take them strictly in that order.  A task of level l may start when every
// task of the levels below it in the same pass has finished (fan-in 0 of a gate comes from anywhere below): each slot has
// one monotonic counter of finished tasks, and the task waits until it reaches  pass * tasks_per_pass + need[seg].
// Everything a task can wait for has a smaller ticket, i.e. was taken by a CTA that is already running — so the scheme
// cannot deadlock whatever the number of resident CTAs — and because the K slots are interleaved level by level, the
// tasks a CTA waits for were handed out K-1 slot-levels earlier: with K >= 3 or so the wait is already over when the
// ticket is taken, and the machine streams through the levels of K tiles without the launch gaps, the half-empty last
// waves and the host launch rate of one kernel per level per tile.
//
// Memory ordering: a task's stores are followed by __syncthreads and a gpu-scope release increment of the slot counter
// by thread 0; a waiting CTA's thread 0 spins on a gpu-scope acquire load and then releases its CTA through
// __syncthreads.  Row loads are strong (relaxed.gpu) loads served by L2: rows are rewritten by other SMs inside this one
// launch (the next pass of a slot reuses the rows), so nothing may be read through a stale L1 line.
//
// Outputs go to up to kMaxDst destinations per tile: a device tensor, page-locked host memory (zero-copy stores over
// PCIe), and — the fused gather of a multi-GPU run — the gathered planes of every peer GPU, mapped through CUDA IPC and
// written over NVLink by the same task that reads the rows (no separate collective, nothing left to overlap).
#pragma once
#include "device_ops.cuh"

namespace b2 {

enum SegType : uint32_t { SEG_STIM = 0, SEG_G2 = 1, SEG_G3 = 2, SEG_WIDE = 3, SEG_OUT = 4 };

struct SweepSeg {
    uint32_t type;
    uint32_t begin, end;  // item range: positions in the input list / gates of the class / positions in the output list
    uint32_t per_task;    // items per task
    uint32_t n_tasks;
    uint32_t task0;       // tasks of this pass in front of this segment
    uint32_t need;        // tasks of this pass that must have finished before a task of this segment starts
    uint32_t pad;
};

static const uint32_t kMaxDst = 10;

struct SweepDst {
    uint64_t *val, *vld;   // planes [n_out][stride]
    uint64_t stride;       // words per output row
    uint64_t word0;        // word of the destination row that takes word 0 of the batch
    uint64_t ring;         // 0, or the destination holds `ring` words per row and is used as a ring
    uint32_t vec;          // 1: 16-byte stores are aligned
    uint32_t pad;
};

struct SweepArgs {
    const uint32_t *fan0, *fan1, *fan2;
    const uint8_t *kind2;
    const uint32_t *wide_off, *wide_in;
    const uint8_t *wide_kind, *wide_nsel, *wide_aux;
    uint64_t *val, *vld;          // K slots back to back, slot_words apart
    uint64_t slot_words;          // n_rows * T
    const SweepSeg *segs;
    const uint32_t *seg_task0;    // task0 of every segment, then tasks_per_pass (for the ticket decode)
    uint32_t n_segs, tasks_per_pass;
    uint32_t K, T, pairs;
    uint32_t n_src, n_fast;
    uint32_t sem;
    uint64_t n_tiles, n_words;    // tiles of this batch; lane-words of this batch (the last tile may be short)
    unsigned long long *ticket;   // one counter
    unsigned long long *done;     // per slot, 16 words (128 B) apart
    uint64_t total_tickets;
    // stimulus
    const uint32_t *in_rows;
    int stim_mode;                // 0 / 1: generated (two-valued / 4-state), 2: planes
    uint64_t seed, word_begin;    // global lane-word of word 0 of the batch (generated stimulus)
    const uint64_t *in_val, *in_vld;
    uint64_t in_stride, in_ring;
    uint32_t in_vec;
    // outputs
    const uint32_t *out_rows;
    uint32_t n_dst;
    SweepDst dst[kMaxDst];
};

__device__ __forceinline__ ulonglong2 ld_coh(const uint64_t *p) {
    ulonglong2 r;
    asm volatile("ld.relaxed.gpu.global.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// One- to three-input gates [g_begin, g_end) of one level, for the slot whose planes are V / M.  The body is k_eval2's
// (engine.cu): explicit phases — indices of all U gates, all plane words, compute and store — every load branch-free, so
// that 4 * U independent 128-bit loads are in flight per thread.
template <int U, int NIN>
__device__ __forceinline__ void sweep_eval(const SweepArgs &a, uint64_t *V, uint64_t *M, uint32_t g_begin, uint32_t g_end) {
    const uint32_t pair = threadIdx.x;
    if (pair >= a.pairs) return;
    const size_t col = (size_t)pair * 2;
    const uint32_t per = blockDim.y * U;
    for (uint32_t base = g_begin; base < g_end; base += per) {
        uint32_t f0[U], f1[U], f2[U], kind[U], orow[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t g = base + threadIdx.y + u * blockDim.y;
            live[u] = g < g_end;
            const uint32_t gg = live[u] ? g : g_end - 1;
            f0[u] = __ldg(a.fan0 + gg);
            f1[u] = __ldg(a.fan1 + gg);
            f2[u] = NIN == 3 ? __ldg(a.fan2 + gg) : 0;
            kind[u] = __ldg(a.kind2 + gg);
            orow[u] = a.n_src + gg;
        }
        ulonglong2 av[U], am[U], bv[U], bm[U], cv[U], cm[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t r0 = (size_t)f0[u] * a.T + col, r1 = (size_t)f1[u] * a.T + col;
            av[u] = ld_coh(V + r0);
            bv[u] = ld_coh(V + r1);
            am[u] = ld_coh(M + r0);
            bm[u] = ld_coh(M + r1);
            if (NIN == 3) {
                const size_t r2 = (size_t)f2[u] * a.T + col;
                cv[u] = ld_coh(V + r2);
                cm[u] = ld_coh(M + r2);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            ulonglong2 rv, rm;
            if (NIN == 3) {
                const uint32_t b = kind[u] >= K_NAND ? kind[u] - 3 : kind[u];  // fold with the positive op, invert once
                gate_word(b, av[u].x, am[u].x, bv[u].x, bm[u].x, rv.x, rm.x, a.sem);
                gate_word(b, av[u].y, am[u].y, bv[u].y, bm[u].y, rv.y, rm.y, a.sem);
                gate_word(b, rv.x, rm.x, cv[u].x, cm[u].x, rv.x, rm.x, a.sem);
                gate_word(b, rv.y, rm.y, cv[u].y, cm[u].y, rv.y, rm.y, a.sem);
                if (kind[u] >= K_NAND) {
                    rv.x = ~rv.x | ~rm.x;
                    rv.y = ~rv.y | ~rm.y;
                }
            } else {
                gate_word(kind[u], av[u].x, am[u].x, bv[u].x, bm[u].x, rv.x, rm.x, a.sem);
                gate_word(kind[u], av[u].y, am[u].y, bv[u].y, bm[u].y, rv.y, rm.y, a.sem);
            }
            if (live[u]) {
                const size_t o = (size_t)orow[u] * a.T + col;
                st_pair(V + o, rv);
                st_pair(M + o, rm);
            }
        }
    }
}

// Stimulus of inputs [begin, end) (positions in the input list) for the tile starting at batch word w0: written straight
// into the source rows of the slot.  Generated words follow b2_fill_stimulus (include/digilogic_b200.h); planes are read
// from device memory or, zero-copy over PCIe, from page-locked host memory.
__device__ __forceinline__ void sweep_stimulus(const SweepArgs &a, uint64_t *V, uint64_t *M, uint32_t begin, uint32_t end, uint64_t w0,
                                               uint32_t tile_words) {
    const uint32_t pair = threadIdx.x;
    if (pair >= a.pairs) return;
    const uint32_t col = pair * 2;
    for (uint32_t i = begin + threadIdx.y; i < end; i += blockDim.y) {
        ulonglong2 v = make_ulonglong2(0, 0), m = make_ulonglong2(0, 0);
        if (a.stim_mode < 2) {
            const uint64_t k = a.seed ^ ((uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull), w = a.word_begin + w0 + col;
            if (col < tile_words) {
                v.x = splitmix64(k ^ (w * 0xD1B54A32D192ED03ull));
                m.x = a.stim_mode == 0 ? ~0ull : (splitmix64(v.x ^ 1) | splitmix64(v.x ^ 2));
            }
            if (col + 1 < tile_words) {
                v.y = splitmix64(k ^ ((w + 1) * 0xD1B54A32D192ED03ull));
                m.y = a.stim_mode == 0 ? ~0ull : (splitmix64(v.y ^ 1) | splitmix64(v.y ^ 2));
            }
        } else {
            uint64_t w = w0 + col;
            if (a.in_ring) w %= a.in_ring;  // tiles never straddle the end of a ring (checked on the host)
            const size_t s = (size_t)i * a.in_stride + w;
            if (a.in_vec && col + 1 < tile_words) {
                v = *(const ulonglong2 *)(a.in_val + s);
                m = *(const ulonglong2 *)(a.in_vld + s);
            } else {
                if (col < tile_words) {
                    v.x = a.in_val[s];
                    m.x = a.in_vld[s];
                }
                if (col + 1 < tile_words) {
                    v.y = a.in_val[s + 1];
                    m.y = a.in_vld[s + 1];
                }
            }
        }
        const size_t o = (size_t)__ldg(a.in_rows + i) * a.T + col;
        st_pair(V + o, v);
        st_pair(M + o, m);
    }
}

// Output rows [begin, end) (positions in the output list) of the tile starting at batch word w0, to every destination.
__device__ __forceinline__ void sweep_outputs(const SweepArgs &a, const uint64_t *V, const uint64_t *M, uint32_t begin, uint32_t end,
                                              uint64_t w0, uint32_t tile_words) {
    const uint32_t pair = threadIdx.x;
    if (pair >= a.pairs) return;
    const uint32_t col = pair * 2;
    if (col >= tile_words) return;
    for (uint32_t i = begin + threadIdx.y; i < end; i += blockDim.y) {
        const size_t o = (size_t)__ldg(a.out_rows + i) * a.T + col;
        const ulonglong2 v = ld_coh(V + o), m = ld_coh(M + o);
#pragma unroll 1
        for (uint32_t d = 0; d < a.n_dst; d++) {
            const SweepDst &t = a.dst[d];
            uint64_t w = t.word0 + w0 + col;
            if (t.ring) w %= t.ring;
            const size_t s = (size_t)i * t.stride + w;
            if (t.vec && col + 1 < tile_words) {
                st_pair(t.val + s, v);
                st_pair(t.vld + s, m);
            } else {
                t.val[s] = v.x;
                t.vld[s] = m.x;
                if (col + 1 < tile_words) {
                    t.val[s + 1] = v.y;
                    t.vld[s + 1] = m.y;
                }
            }
        }
    }
}

template <int U2>
__global__ void __launch_bounds__(256, U2 == 4 ? 3 : 4) k_sweep(const SweepArgs a) {
    __shared__ unsigned long long s_ticket[2];
    const uint32_t tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) s_ticket[0] = atomicAdd(a.ticket, 1ull);
    __syncthreads();
    for (uint32_t it = 0;; it++) {
        const unsigned long long tk = s_ticket[it & 1];
        if (tk >= a.total_tickets) break;
        // the next ticket is taken now; its round trip overlaps this task's work
        unsigned long long next = 0;
        if (tid == 0) next = atomicAdd(a.ticket, 1ull);
        // ---- decode: pass group, segment, slot, index
        const uint64_t per_group = (uint64_t)a.tasks_per_pass * a.K;
        const uint64_t group = tk / per_group;
        const uint32_t r = (uint32_t)(tk - group * per_group);
        const uint32_t q = r / a.K;
        uint32_t lo = 0, hi = a.n_segs;  // last segment with task0 <= q
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(a.seg_task0 + mid) <= q) lo = mid;
            else hi = mid;
        }
        const SweepSeg seg = a.segs[lo];
        const uint32_t rem = r - seg.task0 * a.K;
        const uint32_t slot = rem / seg.n_tasks, idx = rem - slot * seg.n_tasks;
        const uint64_t tile = group * a.K + slot;
        if (tile < a.n_tiles) {
            // ---- wait until the levels below (or, for the stimulus, the slot's previous pass) are complete
            if (tid == 0) {
                const unsigned long long need = group * a.tasks_per_pass + seg.need;
                const unsigned long long *d = a.done + (size_t)slot * 16;
                while (ld_acquire_u64(d) < need) __nanosleep(40);
            }
            __syncthreads();
            uint64_t *V = a.val + (size_t)slot * a.slot_words, *M = a.vld + (size_t)slot * a.slot_words;
            const uint32_t begin = seg.begin + idx * seg.per_task;
            const uint32_t end = min(begin + seg.per_task, seg.end);
            const uint64_t w0 = tile * a.T;
            const uint32_t tile_words = (uint32_t)min((uint64_t)a.T, a.n_words - w0);
            switch (seg.type) {
                case SEG_G2: sweep_eval<U2, 2>(a, V, M, begin, end); break;
                case SEG_G3: sweep_eval<2, 3>(a, V, M, begin, end); break;
                case SEG_WIDE: {
                    WideArgs w{a.wide_off, a.wide_in, a.wide_kind, a.wide_nsel, V, M, V, M, nullptr, begin, end, a.n_src + a.n_fast,
                               a.T, a.pairs, 1, nullptr, nullptr, nullptr, nullptr};
                    w.sem = a.sem;
                    w.aux = a.wide_aux;
                    for (uint32_t g = begin + threadIdx.y; g < end; g += blockDim.y) eval_wide_gate<false>(w, 0, g);
                } break;
                case SEG_STIM: sweep_stimulus(a, V, M, begin, end, w0, tile_words); break;
                default: sweep_outputs(a, V, M, begin, end, w0, tile_words); break;
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                red_release_add_u64(a.done + (size_t)slot * 16, 1ull);
            }
        } else {
            __syncthreads();
        }
        if (tid == 0) s_ticket[(it + 1) & 1] = next;
        __syncthreads();
    }
}

}  // namespace b2
