// engine.cu — sm_100a kernels and the settle loops.
//
// Wire-state store: one row per 1-bit net, `stride` 64-lane words per row, value and valid in
// separate planes, resident in HBM.  A thread owns one 128-bit pair of lane-words of a row; a
// warp therefore reads each fan-in row as one contiguous 512-byte segment (coalesced,
// LDG.128) and the gate function is a handful of LOP3s: this path is HBM/gather bound, there
// is no contraction for the tensor cores to do.
//
// Semantics: with W the wire states and D the external drives, one gsim
// step is W' = G(W), W'[w] = resolve(D[w], F_driver(w)(W)).  run_sim(max_steps) allows
// applications k = 1..N+2 (N = max_steps+1; a cold start spends k = 1 on W_1 = D).  A lane is
// Ok iff W_k == W_{k-1} for some k <= N+2, MaxStepsReached otherwise (state W_{N+1}); a driver
// conflict while computing W_k, k <= N+1, is Err.  For an acyclic single-driver netlist with
// max_steps >= depth the fixed point is unique and is computed level by level in one pass.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>

#include "../../include/digilogic_b200.h"
#include "engine.hpp"
#include "device_ops.cuh"
#include "pair.cuh"

namespace cg = cooperative_groups;

namespace b2 {

// ---------------------------------------------------------------------------- errors / counters
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

void set_last_error(const std::string &msg) { g_last_error = msg; }
const char *get_last_error() { return g_last_error.c_str(); }
uint64_t kernel_launches() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            set_last_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                       \
            return e_ == cudaErrorMemoryAllocation ? B2_ERR_OUT_OF_MEMORY : B2_ERR_CUDA;             \
        }                                                                                             \
    } while (0)

#define LAUNCHED()                                          \
    do {                                                    \
        g_launches.fetch_add(1, std::memory_order_relaxed); \
        CU_TRY(cudaGetLastError());                         \
    } while (0)

#define ST ((cudaStream_t)stream_)
#define CIN ((cudaStream_t)cin_stream_)
#define COUT ((cudaStream_t)cout_stream_)
#define EV(x) ((cudaEvent_t)(x))

// ---------------------------------------------------------------------------- kernels
//
// Common thread mapping: blockDim = (BX, BY), BX * BY = 256.  threadIdx.x walks 128-bit pairs of
// lane-words inside a row, threadIdx.y walks gates; the flat 1-D grid is decomposed as
// (pair block fastest, gate block slowest) so that neighbouring CTAs stream neighbouring
// pieces of the same fan-in rows.

struct Eval2Args {
    const uint32_t *fan0, *fan1, *fan2;  // fan2: three-input gates only (NIN = 3)
    const uint8_t *kind;
    const uint64_t *in_val, *in_vld;  // rows read (A)
    uint64_t *out_val, *out_vld;      // rows written (A in place for the levelised pass, B for unit delay)
    uint64_t *changed;                // [stride] OR of (new ^ old) per lane-word; unit delay only
    uint32_t g_begin, g_end;          // gate range; output row = row_base + gate
    uint32_t row_base;
    uint32_t stride;                  // words per row
    uint32_t pairs;                   // stride / 2
    uint32_t pair_blocks;
    uint8_t *kv;                      // valid-plane elision (KV): row -> 1 if its valid plane is all-ones and not stored
    const uint8_t *chg_prev;          // unit delay: row -> changed in the previous application (nullptr: evaluate every gate)
    uint8_t *chg_cur;                 // unit delay: row -> changed in this application (the frontier of the next one)
    const uint32_t *all_valid;        // KV: 1 if every source row of this run is known valid (then every row is)
    uint32_t sem = 0;                 // SemFlags of the netlist (SURVEY.md A.6-4..7 switches)
    const uint8_t *far = nullptr;     // levelised pass: per gate, bit i = fan-in i was written long ago (L2 evict-first read)
    // levelised pass with discard: two row ranges whose last reader has run (dead rows of the two levels below): dropped from L2
    uint32_t disc_begin[2] = {0, 0}, disc_end[2] = {0, 0};
    // gates [now_begin, now_end) of this launch drive rows that no gate and no output reads: each 128-byte line is dropped
    // from L2 right after the warp has stored it
    uint32_t now_begin = 0, now_end = 0;
    uint32_t pdl = 0;  // levelised pass launched with programmatic stream serialisation: 1 = release dependents early, 2 = late
    // device-side loop: set to 1 when a row written by this application differs from what the buffer held (the state two
    // applications ago) — the period-2 detector of k_unit_delay_loop; nullptr elsewhere
    uint32_t *p2diff = nullptr;
    // k_eval_list: rows read by very many gates (a clock, a reset), staged once per CTA in shared memory (hot_n = 0: none)
    uint32_t hot_n = 0, hot_row[4] = {0, 0, 0, 0};
};

// ---- hot fan-in staging (k_eval_list).  A row with a fan-out in the thousands is re-read from L2 by every listed gate that
// has it as an operand.  The buffer an application reads is constant for the whole launch, so each CTA of the persistent grid
// copies its column block of the (at most four) hottest rows into shared memory once — one bulk asynchronous copy
// (cp.async.bulk, completion on an mbarrier) per row and plane, issued by one thread — and the gate loop takes those operands
// from there.
static constexpr uint32_t kHotRows = 4, kHotPairs = 256, kHotMinFanout = 100;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void stage_hot_rows(const Eval2Args &a, ulonglong2 *s_hot, uint64_t *s_bar, uint32_t pb) {
    const uint32_t tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint32_t p0 = pb * blockDim.x, np = min(blockDim.x, a.pairs - p0), bytes = np * 16;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(s_bar)), "r"(bytes * 2 * a.hot_n) : "memory");
        for (uint32_t h = 0; h < a.hot_n; h++)
            for (uint32_t plane = 0; plane < 2; plane++) {
                const uint64_t *src = (plane ? a.in_vld : a.in_val) + (size_t)a.hot_row[h] * a.stride + (size_t)p0 * 2;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(s_hot + (plane * kHotRows + h) * kHotPairs)),
                             "l"(src), "r"(bytes), "r"(smem_u32(s_bar))
                             : "memory");
            }
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(s_bar)) : "memory");
}
// staged slot of operand `row` of a listed gate, or kHotRows if it is not one of the hot rows
__device__ __forceinline__ uint32_t hot_slot(const Eval2Args &a, uint32_t row) {
    uint32_t s = kHotRows;
#pragma unroll
    for (uint32_t h = 0; h < kHotRows; h++)
        if (h < a.hot_n && row == a.hot_row[h]) s = h;
    return s;
}
// both planes of an operand: a predicated-off global load (no branch, so the loads of all operands of an iteration are in flight
// together) patched from the staged copy when the row is hot
__device__ __forceinline__ void ld_operand(const Eval2Args &a, const ulonglong2 *s_hot, uint32_t row, size_t off, ulonglong2 &v, ulonglong2 &m) {
    const uint32_t s = s_hot ? hot_slot(a, row) : kHotRows;
    const bool staged = s < kHotRows;
    asm volatile("{ .reg .pred q; setp.eq.u32 q, %3, 0; @q ld.volatile.global.v2.u64 {%0,%1}, [%2]; }" : "=l"(v.x), "=l"(v.y) : "l"(a.in_val + off), "r"((uint32_t)staged) : "memory");
    asm volatile("{ .reg .pred q; setp.eq.u32 q, %3, 0; @q ld.volatile.global.v2.u64 {%0,%1}, [%2]; }" : "=l"(m.x), "=l"(m.y) : "l"(a.in_vld + off), "r"((uint32_t)staged) : "memory");
    if (staged) {
        v = s_hot[s * kHotPairs + threadIdx.x];
        m = s_hot[(kHotRows + s) * kHotPairs + threadIdx.x];
    }
}

// Changed-row frontier write: with a whole warp on one gate, one ballot and one byte store.
__device__ __forceinline__ void mark_changed(uint8_t *chg, uint32_t row, bool changed) {
    if (blockDim.x >= 32) {
        // lanes past the end of a short row have returned: vote among the ones still here (lane 0 always is)
        const unsigned any = __ballot_sync(__activemask(), changed);
        if (any && (threadIdx.x & 31) == 0) chg[row] = 1;
    } else if (changed) {
        chg[row] = 1;
    }
}

// One- and two-input gates.  U gates per thread keep 4*U independent 128-bit loads in flight.
// KV (levelised only): a row whose fan-ins are all "known valid" for this tile has an all-ones
// valid plane by construction (every gate maps valid inputs to a valid output), so that plane is
// neither read nor written: 24 B instead of 48 B per gate-word.  Rows that may carry X/Z take the
// full two-plane path; the result is bit-identical either way.
// NIN = 3: three-input gates, ((a op b) op c) with one inversion at the end for NAND/NOR/XNOR.
template <int U, bool JACOBI, bool KV, int NIN = 2>
__global__ void __launch_bounds__(256, JACOBI ? 2 : ((U == 2 && NIN == 2) ? (KV ? 4 : 5) : 3)) k_eval2(const Eval2Args a) {
    const uint32_t pb = blockIdx.x % a.pair_blocks, gb = blockIdx.x / a.pair_blocks;
    const uint32_t pair = pb * blockDim.x + threadIdx.x;
    // Programmatic dependent launch (levelised pass, a.pdl): the next level's kernel of this stream may be launched while
    // this one runs — early: as soon as every CTA of this grid has started; late: when every CTA is about to finish — and
    // overlaps its launch and its index loads (which depend on nothing this grid writes) with this grid's tail; it does not
    // touch a row before griddepcontrol.wait below, i.e. before this grid has completed and its writes are visible.
    if (!JACOBI && a.pdl == 1) asm volatile("griddepcontrol.launch_dependents;");
    if (pair >= a.pairs) return;
    const size_t col = (size_t)pair * 2;
    const uint32_t g0 = a.g_begin + gb * (blockDim.y * U) + threadIdx.y;

    // Explicit phases, each a fully unrolled loop over the U gates of this thread, so that every
    // dependent step (indices -> elision flags -> plane words) costs one memory round trip for all
    // U gates together; loads are branch-free (tail gates are clamped, elided planes predicated off).
    uint32_t f0[U], f1[U], f2[U], kind[U], orow[U], far[U];
    bool live[U];
    uint64_t pol_first = 0;
    if (!JACOBI && !KV) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint32_t g = g0 + u * blockDim.y;
        live[u] = g < a.g_end;
        const uint32_t gg = live[u] ? g : a.g_end - 1;
        f0[u] = a.fan0[gg];
        f1[u] = a.fan1[gg];
        f2[u] = NIN == 3 ? a.fan2[gg] : 0;
        kind[u] = a.kind[gg];
        orow[u] = a.row_base + gg;
        if (!JACOBI && !KV) far[u] = a.far ? a.far[gg] : 0;
    }
    if (!JACOBI && a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!JACOBI) {
        // Rows nobody will read again: invalidate their L2 lines without write-back, so they never cost HBM bandwidth (the level
        // that read them last ran in an earlier launch on this stream).  One 128-byte line per plane per thread and round.
#pragma unroll
        for (int r = 0; r < 2; r++)
            if (a.disc_end[r] > a.disc_begin[r]) {
                // (by the CTAs that start first: the sooner a dead line is dropped, the likelier it has not been written back yet)
                const uint64_t lines = (uint64_t)(a.disc_end[r] - a.disc_begin[r]) * (a.stride >> 4);
                const uint32_t cta_threads = blockDim.x * blockDim.y;
                const uint32_t n_cta = min(gridDim.x, 296u * (256u / cta_threads));
                if (blockIdx.x >= n_cta) break;
                const uint64_t nt = (uint64_t)n_cta * cta_threads;
                for (uint64_t i = (uint64_t)blockIdx.x * cta_threads + threadIdx.y * blockDim.x + threadIdx.x; i < lines; i += nt) {
                    const size_t w = (size_t)a.disc_begin[r] * a.stride + i * 16;
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(a.out_val + w) : "memory");
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(a.out_vld + w) : "memory");
                }
            }
    }

    bool k0[U], k1[U], k2[U], act[U];
    bool any_act = false;
    const bool allv = KV && *a.all_valid != 0;  // uniform: with all drives valid no per-row flag is read or written
#pragma unroll
    for (int u = 0; u < U; u++) {
        k0[u] = KV && (allv || a.kv[f0[u]] != 0);
        k1[u] = KV && (allv || a.kv[f1[u]] != 0);  // NOT gates have f1 == f0
        k2[u] = NIN == 3 ? (KV && (allv || a.kv[f2[u]] != 0)) : true;
        // unit delay: a gate whose fan-ins did not change in the previous application keeps its
        // output; it is evaluated only if that output row itself changed (the other buffer is stale)
        act[u] = !JACOBI || !a.chg_prev ||
                 (a.chg_prev[f0[u]] | a.chg_prev[f1[u]] | (NIN == 3 ? a.chg_prev[f2[u]] : 0) | a.chg_prev[orow[u]]) != 0;
        any_act |= act[u] && live[u];
    }
    if (JACOBI && !any_act) return;
    ulonglong2 av[U], am[U], bv[U], bm[U], cv[U], cm[U], ov[U], om[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const size_t r0 = (size_t)f0[u] * a.stride + col, r1 = (size_t)f1[u] * a.stride + col;
        if (NIN == 3) {
            const size_t r2 = (size_t)f2[u] * a.stride + col;
            const bool skip = JACOBI && !act[u];
            cv[u] = ld_stream_unless(a.in_val + r2, skip);
            cm[u] = ld_stream_unless(a.in_vld + r2, skip || (KV && k2[u]));
        }
        if (JACOBI) {
            const bool skip = !act[u];
            const size_t o = (size_t)orow[u] * a.stride + col;
            av[u] = ld_stream_unless(a.in_val + r0, skip);
            bv[u] = ld_stream_unless(a.in_val + r1, skip);
            am[u] = ld_stream_unless(a.in_vld + r0, skip);
            bm[u] = ld_stream_unless(a.in_vld + r1, skip);
            ov[u] = ld_stream_unless(a.in_val + o, skip);
            om[u] = ld_stream_unless(a.in_vld + o, skip);
        } else {
            if (KV) {
                av[u] = ld_dep(a.in_val + r0);
                bv[u] = ld_dep(a.in_val + r1);
            } else {
                // (ld_dep: with programmatic dependent launch the rows are written while this grid is resident — not .nc data)
                av[u] = ld_dep_hint(a.in_val + r0, far[u] & 1, pol_first);
                bv[u] = ld_dep_hint(a.in_val + r1, far[u] & 2, pol_first);
            }
            if (KV) {
                am[u] = ld_dep_unless(a.in_vld + r0, k0[u]);
                bm[u] = ld_dep_unless(a.in_vld + r1, k1[u]);
            } else {
                am[u] = ld_dep_hint(a.in_vld + r0, far[u] & 1, pol_first);
                bm[u] = ld_dep_hint(a.in_vld + r1, far[u] & 2, pol_first);
            }
        }
    }
    uint64_t dx = 0, dy = 0;
#pragma unroll
    for (int u = 0; u < U; u++) {
        ulonglong2 rv, rm;
        if (NIN == 3) {
            const uint32_t base = kind[u] >= K_NAND ? kind[u] - 3 : kind[u];  // fold with the positive op
            gate_word(base, av[u].x, am[u].x, bv[u].x, bm[u].x, rv.x, rm.x, a.sem);
            gate_word(base, av[u].y, am[u].y, bv[u].y, bm[u].y, rv.y, rm.y, a.sem);
            gate_word(base, rv.x, rm.x, cv[u].x, cm[u].x, rv.x, rm.x, a.sem);
            gate_word(base, rv.y, rm.y, cv[u].y, cm[u].y, rv.y, rm.y, a.sem);
            if (kind[u] >= K_NAND) {
                rv.x = ~rv.x | ~rm.x;
                rv.y = ~rv.y | ~rm.y;
            }
        } else {
            gate_word(kind[u], av[u].x, am[u].x, bv[u].x, bm[u].x, rv.x, rm.x, a.sem);
            gate_word(kind[u], av[u].y, am[u].y, bv[u].y, bm[u].y, rv.y, rm.y, a.sem);
        }
        bool ch = false;
        if (live[u] && act[u]) {
            const size_t o = (size_t)orow[u] * a.stride + col;
            const bool okv = KV && k0[u] && k1[u] && k2[u];
            st_pair(a.out_val + o, rv);
            if (!okv) st_pair(a.out_vld + o, rm);
            if (!JACOBI && (pair & 7) == 0 && orow[u] - a.row_base >= a.now_begin && orow[u] - a.row_base < a.now_end) {
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(a.out_val + o) : "memory");
                if (!okv) asm volatile("discard.global.L2 [%0], 128;" ::"l"(a.out_vld + o) : "memory");
            }
            if (KV && !allv && pb == 0 && threadIdx.x == 0) a.kv[orow[u]] = okv ? 1 : 0;
            if (JACOBI) {
                const uint64_t ex = (rv.x ^ ov[u].x) | (rm.x ^ om[u].x), ey = (rv.y ^ ov[u].y) | (rm.y ^ om[u].y);
                if (kind[u] != K_RAW) {  // a register's clock history is internal state: it keeps no lane alive
                    dx |= ex;
                    dy |= ey;
                }
                ch = (ex | ey) != 0;
            }
        }
        if (JACOBI && a.chg_cur) mark_changed(a.chg_cur, orow[u], ch);
    }
    if (JACOBI) {
        if (dx) atomicOr((unsigned long long *)&a.changed[col], (unsigned long long)dx);
        if (dy) atomicOr((unsigned long long *)&a.changed[col + 1], (unsigned long long)dy);
    }
    if (!JACOBI && a.pdl == 2) asm volatile("griddepcontrol.launch_dependents;");
}

// ---- unit-delay frontier as a compacted work list (gsim's component_update_queue, at row
// granularity).  k_frontier_scan: one thread per fast gate decides whether the gate has to be evaluated in
// this application — a fan-in row or its own row changed in the previous one — and appends it to the list
// of its class with one warp-aggregated atomic.  k_eval_list then evaluates exactly the listed gates with
// a persistent grid, so an application in which little changes costs a few microseconds instead of a
// sweep of CTAs over every gate.
// Changed-row flags are written by other CTAs earlier in the same launch when the scan runs inside the device-side
// loop: read them past L1 (volatile), like the rows themselves.
__device__ __forceinline__ uint32_t ld_flag(const uint8_t *p) { return *(const volatile uint8_t *)p; }

__device__ __forceinline__ void frontier_scan_body(const uint32_t *fan0, const uint32_t *fan1, const uint32_t *fan2, uint32_t n_g2,
                                                   uint32_t n_fast, uint32_t row_base, const uint8_t *kind, const uint8_t *chg_prev,
                                                   uint4 *list2, uint4 *list3, uint32_t *list_copy, uint32_t *counts, uint32_t tid,
                                                   uint32_t g) {
    bool act = false, refresh = false;
    uint4 e = make_uint4(0, 0, 0, 0);  // {gate | kind << 28, fan-in rows}: the evaluation needs no further index loads
    if (g < n_fast) {
        e.y = fan0[g];
        e.z = fan1[g];
        e.w = g >= n_g2 ? fan2[g] : e.y;
        e.x = g | ((uint32_t)kind[g] << 28);
        if (!chg_prev) {
            act = true;  // flags not valid (first application after a reset): every gate is evaluated
        } else {
            act = (ld_flag(chg_prev + e.y) | ld_flag(chg_prev + e.z)) != 0;
            if (g >= n_g2) act |= ld_flag(chg_prev + e.w) != 0;
            // no fan-in changed but the gate's own row did: its value is final, only the other buffer's copy is stale
            refresh = !act && ld_flag(chg_prev + row_base + g) != 0;
        }
    }
    const uint32_t lane = tid & 31, lt = (1u << lane) - 1;
    const unsigned m2 = __ballot_sync(0xFFFFFFFFu, act && g < n_g2), m3 = __ballot_sync(0xFFFFFFFFu, act && g >= n_g2),
                   mc = __ballot_sync(0xFFFFFFFFu, refresh);
    uint32_t b2 = 0, b3 = 0, bc = 0;
    if (lane == 0) {
        if (m2) b2 = atomicAdd(&counts[0], (uint32_t)__popc(m2));
        if (m3) b3 = atomicAdd(&counts[1], (uint32_t)__popc(m3));
        if (mc) bc = atomicAdd(&counts[2], (uint32_t)__popc(mc));
    }
    b2 = __shfl_sync(0xFFFFFFFFu, b2, 0);
    b3 = __shfl_sync(0xFFFFFFFFu, b3, 0);
    bc = __shfl_sync(0xFFFFFFFFu, bc, 0);
    if (act) {
        if (g < n_g2) list2[b2 + __popc(m2 & lt)] = e;
        else list3[b3 + __popc(m3 & lt)] = e;
    } else if (refresh) {
        list_copy[bc + __popc(mc & lt)] = row_base + g;
    }
}

__global__ void __launch_bounds__(256) k_frontier_scan(const uint32_t *fan0, const uint32_t *fan1, const uint32_t *fan2, uint32_t n_g2,
                                                       uint32_t n_fast, uint32_t row_base, const uint8_t *kind, const uint8_t *chg_prev, uint4 *list2,
                                                       uint4 *list3, uint32_t *list_copy, uint32_t *counts) {
    frontier_scan_body(fan0, fan1, fan2, n_g2, n_fast, row_base, kind, chg_prev, list2, list3, list_copy, counts, threadIdx.x,
                       blockIdx.x * blockDim.x + threadIdx.x);
}

// Rows whose value did not change in this application but did in the previous one: B[row] = A[row] brings the
// other buffer up to date (32 B per row-word instead of a full re-evaluation).
__device__ __forceinline__ void copy_list_body(const uint32_t *rows, const uint32_t *count, const uint64_t *a_val,
                                               const uint64_t *a_vld, uint64_t *b_val, uint64_t *b_vld, uint32_t stride, uint32_t pairs,
                                               uint32_t pair_blocks, uint32_t bid, uint32_t nblk, uint32_t *p2diff = nullptr) {
    const uint32_t n = *count;
    const uint32_t pb = bid % pair_blocks, rbi = bid / pair_blocks, rbs = nblk / pair_blocks;
    const uint32_t pair = pb * blockDim.x + threadIdx.x;
    if (pair >= pairs) return;
    for (uint32_t i = rbi * blockDim.y + threadIdx.y; i < n; i += rbs * blockDim.y) {
        const size_t o = (size_t)rows[i] * stride + (size_t)pair * 2;
        const ulonglong2 v = ld_row(a_val + o), m = ld_row(a_vld + o);
        if (p2diff) {
            const ulonglong2 pv = ld_row(b_val + o), pm = ld_row(b_vld + o);
            if ((v.x ^ pv.x) | (m.x ^ pm.x) | (v.y ^ pv.y) | (m.y ^ pm.y)) *p2diff = 1;
        }
        st_pair(b_val + o, v);
        st_pair(b_vld + o, m);
    }
}

__global__ void __launch_bounds__(256) k_copy_list(const uint32_t *rows, const uint32_t *count, const uint64_t *a_val,
                                                   const uint64_t *a_vld, uint64_t *b_val, uint64_t *b_vld, uint32_t stride,
                                                   uint32_t pairs, uint32_t pair_blocks) {
    copy_list_body(rows, count, a_val, a_vld, b_val, b_vld, stride, pairs, pair_blocks, blockIdx.x, gridDim.x);
}

// Gates of `list[0 .. *count)` (one/two-input, or three-input with NIN = 3): B[row] = F(A), change detection
// against A[row], changed-row flag.  Persistent: the grid is sized for the machine, CTAs stride over the list.
template <int NIN, bool HOT = false>
__device__ __forceinline__ void eval_list_body(const Eval2Args &a, const uint4 *list, const uint32_t *count, uint32_t bid, uint32_t nblk,
                                               const ulonglong2 *s_hot = nullptr) {
    constexpr int U = 2;
    const uint32_t n = *count;
    const uint32_t pb = bid % a.pair_blocks, gbi = bid / a.pair_blocks, gbs = nblk / a.pair_blocks;
    const uint32_t pair = pb * blockDim.x + threadIdx.x;
    if (pair >= a.pairs || n == 0) return;
    const size_t col = (size_t)pair * 2;
    const uint32_t per = blockDim.y * U;
    uint64_t dx = 0, dy = 0, d2 = 0;
    // list entries of the next iteration are fetched while the plane words of this one are in flight
    uint4 nxt[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint32_t idx = gbi * per + threadIdx.y + u * blockDim.y;
        nxt[u] = list[idx < n ? idx : n - 1];
    }
    for (uint32_t base = gbi * per; base < n; base += gbs * per) {
        uint32_t f0[U], f1[U], f2[U], kind[U], orow[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            live[u] = base + threadIdx.y + u * blockDim.y < n;
            f0[u] = nxt[u].y;
            f1[u] = nxt[u].z;
            f2[u] = nxt[u].w;
            kind[u] = nxt[u].x >> 28;
            orow[u] = a.row_base + (nxt[u].x & 0x0FFFFFFFu);
            const uint32_t idx = base + gbs * per + threadIdx.y + u * blockDim.y;
            nxt[u] = list[idx < n ? idx : n - 1];
        }
        ulonglong2 av[U], am[U], bv[U], bm[U], cv[U], cm[U], ov[U], om[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t r0 = (size_t)f0[u] * a.stride + col, r1 = (size_t)f1[u] * a.stride + col, o = (size_t)orow[u] * a.stride + col;
            if (HOT) {
                ld_operand(a, s_hot, f0[u], r0, av[u], am[u]);
                ld_operand(a, s_hot, f1[u], r1, bv[u], bm[u]);
                if (NIN == 3) ld_operand(a, s_hot, f2[u], (size_t)f2[u] * a.stride + col, cv[u], cm[u]);
            } else {
                av[u] = ld_row(a.in_val + r0);
                bv[u] = ld_row(a.in_val + r1);
                am[u] = ld_row(a.in_vld + r0);
                bm[u] = ld_row(a.in_vld + r1);
                if (NIN == 3) {
                    const size_t r2 = (size_t)f2[u] * a.stride + col;
                    cv[u] = ld_row(a.in_val + r2);
                    cm[u] = ld_row(a.in_vld + r2);
                }
            }
            ov[u] = ld_row(a.in_val + o);
            om[u] = ld_row(a.in_vld + o);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            ulonglong2 rv, rm;
            if (NIN == 3) {
                const uint32_t base_kind = kind[u] >= K_NAND ? kind[u] - 3 : kind[u];
                gate_word(base_kind, av[u].x, am[u].x, bv[u].x, bm[u].x, rv.x, rm.x, a.sem);
                gate_word(base_kind, av[u].y, am[u].y, bv[u].y, bm[u].y, rv.y, rm.y, a.sem);
                gate_word(base_kind, rv.x, rm.x, cv[u].x, cm[u].x, rv.x, rm.x, a.sem);
                gate_word(base_kind, rv.y, rm.y, cv[u].y, cm[u].y, rv.y, rm.y, a.sem);
                if (kind[u] >= K_NAND) {
                    rv.x = ~rv.x | ~rm.x;
                    rv.y = ~rv.y | ~rm.y;
                }
            } else {
                gate_word(kind[u], av[u].x, am[u].x, bv[u].x, bm[u].x, rv.x, rm.x, a.sem);
                gate_word(kind[u], av[u].y, am[u].y, bv[u].y, bm[u].y, rv.y, rm.y, a.sem);
            }
            bool ch = false;
            if (live[u]) {
                const size_t o = (size_t)orow[u] * a.stride + col;
                if (a.p2diff) {  // what this buffer held = the row two applications ago
                    const ulonglong2 pv = ld_row(a.out_val + o), pm = ld_row(a.out_vld + o);
                    d2 |= (rv.x ^ pv.x) | (rm.x ^ pm.x) | (rv.y ^ pv.y) | (rm.y ^ pm.y);
                }
                st_pair(a.out_val + o, rv);
                st_pair(a.out_vld + o, rm);
                const uint64_t ex = (rv.x ^ ov[u].x) | (rm.x ^ om[u].x), ey = (rv.y ^ ov[u].y) | (rm.y ^ om[u].y);
                if (kind[u] != K_RAW) {  // a register's clock history is internal state: it keeps no lane alive
                    dx |= ex;
                    dy |= ey;
                }
                ch = (ex | ey) != 0;
            }
            if (blockDim.x >= 32) {  // a whole warp on one gate: one ballot, one byte store
                const unsigned any = __ballot_sync(__activemask(), ch);
                if (any && (threadIdx.x & 31) == 0) a.chg_cur[orow[u]] = 1;
            } else if (ch) {
                a.chg_cur[orow[u]] = 1;
            }
        }
    }
    if (dx) atomicOr((unsigned long long *)&a.changed[col], (unsigned long long)dx);
    if (dy) atomicOr((unsigned long long *)&a.changed[col + 1], (unsigned long long)dy);
    if (d2) *a.p2diff = 1;
}

template <int NIN>
__global__ void __launch_bounds__(256, NIN == 3 ? 2 : 3) k_eval_list(const Eval2Args a, const uint4 *list, const uint32_t *count) {
    eval_list_body<NIN>(a, list, count, blockIdx.x, gridDim.x);
}
// the same with the hot rows of a.hot_row staged in shared memory (B2_OPT_HOT_STAGING)
template <int NIN>
__global__ void __launch_bounds__(256, NIN == 3 ? 2 : 3) k_eval_list_hot(const Eval2Args a, const uint4 *list, const uint32_t *count) {
    __shared__ __align__(128) ulonglong2 s_hot[2 * kHotRows * kHotPairs];  // [plane][hot row][pair of this CTA's column block]
    __shared__ __align__(8) uint64_t s_bar;
    if (*count == 0) return;  // uniform over the grid; an idle launch stages nothing
    stage_hot_rows(a, s_hot, &s_bar, blockIdx.x % a.pair_blocks);
    eval_list_body<NIN, true>(a, list, count, blockIdx.x, gridDim.x, s_hot);
}

template <bool JACOBI>
__device__ __forceinline__ void eval_wide_body(const WideArgs &a, uint32_t pb, uint32_t gb) {
    eval_wide_gate<JACOBI>(a, pb, a.g_begin + gb * blockDim.y + threadIdx.y);
}

// wide gates named by a work list (unit-delay frontier): the scan has already decided that they have to be evaluated
__device__ __forceinline__ void eval_wide_list_body(WideArgs a, const uint32_t *list, const uint32_t *count, uint32_t bid,
                                                    uint32_t nblk) {
    const uint32_t n = *count;
    const uint32_t pb = bid % a.pair_blocks, gbi = bid / a.pair_blocks, gbs = nblk / a.pair_blocks;
    a.chg_prev = nullptr;
    for (uint32_t i = gbi * blockDim.y + threadIdx.y; i < n; i += gbs * blockDim.y) eval_wide_gate<true>(a, pb, list[i]);
}

template <bool JACOBI>
__global__ void __launch_bounds__(256) k_eval_wide(const WideArgs a) {
    eval_wide_body<JACOBI>(a, blockIdx.x % a.pair_blocks, blockIdx.x / a.pair_blocks);
}

__global__ void __launch_bounds__(256) k_eval_wide_list(const WideArgs a, const uint32_t *list, const uint32_t *count) {
    eval_wide_list_body(a, list, count, blockIdx.x, gridDim.x);
}

// Frontier scan of the wide gates (n-ary gates, mux slices, registers, adder bits): listed when an input row or the
// gate's own row changed in the previous application (chg_prev == nullptr: all of them).
__device__ __forceinline__ void frontier_scan_wide_body(const uint32_t *off, const uint32_t *in, uint32_t n_wide, uint32_t row_base,
                                                        const uint8_t *chg_prev, uint32_t *list, uint32_t *count, uint32_t tid,
                                                        uint32_t g) {
    bool act = false;
    if (g < n_wide) {
        if (!chg_prev) {
            act = true;
        } else {
            uint32_t any = ld_flag(chg_prev + row_base + g);
            for (uint32_t i = off[g]; i < off[g + 1] && !any; i++) any |= ld_flag(chg_prev + in[i]);
            act = any != 0;
        }
    }
    const uint32_t lane = tid & 31;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, act);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (act) list[base + __popc(m & ((1u << lane) - 1))] = g;
}

__global__ void __launch_bounds__(256) k_frontier_scan_wide(const uint32_t *off, const uint32_t *in, uint32_t n_wide, uint32_t row_base,
                                                            const uint8_t *chg_prev, uint32_t *list, uint32_t *count) {
    frontier_scan_wide_body(off, in, n_wide, row_base, chg_prev, list, count, threadIdx.x, blockIdx.x * blockDim.x + threadIdx.x);
}

// Source rows take their external drive: out[row] = D[row] (rows [0, n_rows) of both stores).
// kv (nullable, pre-set to 1 per source row): cleared when any in-use lane-word of the row's
// valid plane is not all-ones.
template <bool JACOBI>
__device__ __forceinline__ void apply_drive_body(const uint64_t *dval, const uint64_t *dvld, const uint64_t *old_val,
                                                 const uint64_t *old_vld, uint64_t *out_val, uint64_t *out_vld, uint64_t *changed,
                                                 uint32_t n_rows, uint32_t stride, uint32_t pairs, uint8_t *kv, uint32_t words_in_use,
                                                 uint8_t *chg_cur, uint32_t pb, uint32_t rb) {
    const uint32_t pair = pb * blockDim.x + threadIdx.x;
    const uint32_t row = rb * blockDim.y + threadIdx.y;
    if (pair >= pairs || row >= n_rows) return;
    const size_t col = (size_t)pair * 2, o = (size_t)row * stride + col;
    const ulonglong2 v = ld_row(dval + o), m = ld_row(dvld + o);
    if (JACOBI) {
        const ulonglong2 ov = ld_row(old_val + o), om = ld_row(old_vld + o);
        const uint64_t dx = (v.x ^ ov.x) | (m.x ^ om.x), dy = (v.y ^ ov.y) | (m.y ^ om.y);
        if (dx) atomicOr((unsigned long long *)&changed[col], (unsigned long long)dx);
        if (dy) atomicOr((unsigned long long *)&changed[col + 1], (unsigned long long)dy);
        if (chg_cur && (dx | dy)) chg_cur[row] = 1;
    }
    st_pair(out_val + o, v);
    st_pair(out_vld + o, m);
    if (kv) {
        const bool bad = (col < words_in_use && m.x != ~0ull) || (col + 1 < words_in_use && m.y != ~0ull);
        if (bad) kv[row] = 0;
    }
}

template <bool JACOBI>
__global__ void __launch_bounds__(256) k_apply_drive(const uint64_t *dval, const uint64_t *dvld, const uint64_t *old_val,
                                                     const uint64_t *old_vld, uint64_t *out_val, uint64_t *out_vld,
                                                     uint64_t *changed, uint32_t n_rows, uint32_t stride, uint32_t pairs,
                                                     uint32_t pair_blocks, uint8_t *kv = nullptr, uint32_t words_in_use = 0,
                                                     uint8_t *chg_cur = nullptr) {
    apply_drive_body<JACOBI>(dval, dvld, old_val, old_vld, out_val, out_vld, changed, n_rows, stride, pairs, kv, words_in_use, chg_cur,
                             blockIdx.x % pair_blocks, blockIdx.x / pair_blocks);
}

// all_valid = AND of the source rows' flags (one CTA)
__global__ void __launch_bounds__(256) k_all_valid(const uint8_t *kv, uint32_t n_src, uint32_t *all_valid) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_src; i += blockDim.x)
        if (!kv[i]) bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) *all_valid = bad ? 0u : 1u;
}

// expands elided valid planes back to all-ones words (leaving the store plain two-plane)
__global__ void __launch_bounds__(256) k_materialise_valid(uint64_t *vld, uint8_t *kv, const uint32_t *all_valid, uint32_t n_rows,
                                                           uint32_t stride, uint32_t pairs, uint32_t pair_blocks) {
    const uint32_t pb = blockIdx.x % pair_blocks, rb = blockIdx.x / pair_blocks;
    const uint32_t pair = pb * blockDim.x + threadIdx.x;
    const uint32_t row = rb * blockDim.y + threadIdx.y;
    if (pair >= pairs || row >= n_rows) return;
    if (*all_valid || kv[row]) st_pair(vld + (size_t)row * stride + (size_t)pair * 2, make_ulonglong2(~0ull, ~0ull));
}

// External drives on gate-driven rows: resolve(D, gate output) and conflict detection
// (both operands not High-Z).  mode 0: merge + conflict; mode 1 (cold start,
// component outputs still High-Z): the row simply takes the drive.
__global__ void __launch_bounds__(256) k_extra_drive(const uint32_t *rows, uint32_t n, const uint64_t *xval,
                                                     const uint64_t *xvld, uint64_t *val, uint64_t *vld, uint64_t *conf,
                                                     uint32_t stride, int mode, int lenient = 0) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= stride) return;
    uint64_t c = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (rows[i] == 0xFFFFFFFFu) continue;  // slot of a tri-state net: k_resolve merges it
        const size_t o = (size_t)rows[i] * stride + w, x = (size_t)i * stride + w;
        const uint64_t dv = xval[x], dm = xvld[x];
        if (mode == 1) {
            val[o] = dv;
            vld[o] = dm;
        } else {
            const uint64_t gv = val[o], gm = vld[o];
            c |= conflict_of(dv, dm, gv, gm, lenient);
            val[o] = gv | dv;
            vld[o] = gm | dm;
        }
    }
    if (mode == 0 && c) atomicOr((unsigned long long *)&conf[w], (unsigned long long)c);
}

// Tri-state nets (SURVEY.md A.4 wire update, restated for nets whose drivers can be High-Z):
//   W'[net] = resolve(D[net], outputs of its drivers),  resolve = plane-wise OR,
//   conflict where two operands are not High-Z.
// The drivers' outputs were written to their own rows of B by the gate kernels of this application;
// this kernel runs right after them, so the wire sees its drivers with no extra delay.  Every entry is
// evaluated in every application (such nets are few); `cold` = first step of a fresh simulator, when
// all component outputs are still High-Z and the wire is just its drive.
struct ResolveArgs {
    const uint32_t *off, *in;
    const int32_t *xslot;             // entry -> slot of its external drive in the extra store, or -1
    const uint64_t *xval, *xvld;      // extra drive store [slot][stride]
    const uint64_t *a_val, *a_vld;    // previous state (change detection)
    uint64_t *b_val, *b_vld;          // driver rows are read from, the wire row is written to, this buffer
    uint64_t *changed, *conf;
    uint8_t *chg_cur;
    uint32_t n_res, row_base, stride, pairs, pair_blocks;
    int cold;
    int lenient;  // A.6-2 alternative: agreeing valid drivers do not conflict
    uint32_t *p2diff = nullptr;  // as in Eval2Args
};

__device__ __forceinline__ void resolve_body(const ResolveArgs &a, uint32_t pb, uint32_t rb) {
    const uint32_t pair = pb * blockDim.x + threadIdx.x;
    const uint32_t r = rb * blockDim.y + threadIdx.y;
    if (pair >= a.pairs || r >= a.n_res) return;
    const size_t col = (size_t)pair * 2;
    ulonglong2 v = make_ulonglong2(0, 0), m = make_ulonglong2(0, 0);
    const int32_t slot = a.xslot[r];
    if (slot >= 0) {
        const size_t x = (size_t)slot * a.stride + col;
        v = *(const ulonglong2 *)(a.xval + x);
        m = *(const ulonglong2 *)(a.xvld + x);
    }
    uint64_t cx = 0, cy = 0;
    if (!a.cold)
        for (uint32_t i = a.off[r]; i < a.off[r + 1]; i++) {
            const size_t o = (size_t)a.in[i] * a.stride + col;
            const ulonglong2 ov = ld_row(a.b_val + o), om = ld_row(a.b_vld + o);  // written by other CTAs earlier in this launch
            cx |= conflict_of(v.x, m.x, ov.x, om.x, a.lenient);
            cy |= conflict_of(v.y, m.y, ov.y, om.y, a.lenient);
            v.x |= ov.x; v.y |= ov.y;
            m.x |= om.x; m.y |= om.y;
        }
    const size_t o = (size_t)(a.row_base + r) * a.stride + col;
    if (a.changed) {
        const ulonglong2 pv = ld_row(a.a_val + o), pm = ld_row(a.a_vld + o);
        const uint64_t dx = (v.x ^ pv.x) | (m.x ^ pm.x), dy = (v.y ^ pv.y) | (m.y ^ pm.y);
        if (dx) atomicOr((unsigned long long *)&a.changed[col], (unsigned long long)dx);
        if (dy) atomicOr((unsigned long long *)&a.changed[col + 1], (unsigned long long)dy);
        if (a.chg_cur && (dx | dy)) a.chg_cur[a.row_base + r] = 1;
    }
    if (cx) atomicOr((unsigned long long *)&a.conf[col], (unsigned long long)cx);
    if (cy) atomicOr((unsigned long long *)&a.conf[col + 1], (unsigned long long)cy);
    if (a.p2diff) {
        const ulonglong2 pv = ld_row(a.b_val + o), pm = ld_row(a.b_vld + o);
        if ((v.x ^ pv.x) | (m.x ^ pm.x) | (v.y ^ pv.y) | (m.y ^ pm.y)) *a.p2diff = 1;
    }
    st_pair(a.b_val + o, v);
    st_pair(a.b_vld + o, m);
}

__global__ void __launch_bounds__(256) k_resolve(const ResolveArgs a) {
    resolve_body(a, blockIdx.x % a.pair_blocks, blockIdx.x / a.pair_blocks);
}

// Per-lane bookkeeping after one application:
//   conflict |= conf_step & live (only while k < last); live &= changed & ~conflict;
//   on the last (check-only) application: unsettled = live.
__device__ __forceinline__ void status_step_body(uint64_t *changed, uint64_t *conf_step, uint64_t *live, uint64_t *conflict,
                                                 uint64_t *unsettled, RunFlags *flags, uint32_t words, uint64_t n_lanes,
                                                 int count_conflicts, int is_last, uint32_t w, uint32_t lane) {
    uint64_t lv = 0, cf = 0, un = 0;
    if (w < words) {
        const uint64_t mask = lane_mask_of(w, n_lanes);
        lv = live[w] & mask;
        cf = conflict[w];
        if (count_conflicts) cf |= conf_step[w] & lv;
        lv &= changed[w] & ~cf;
        live[w] = lv;
        conflict[w] = cf;
        if (is_last) {
            un = lv;
            unsettled[w] = un;
        }
        changed[w] = 0;
        conf_step[w] = 0;
    }
    const unsigned any_l = __ballot_sync(0xFFFFFFFFu, lv != 0), any_c = __ballot_sync(0xFFFFFFFFu, cf != 0),
                   any_u = __ballot_sync(0xFFFFFFFFu, un != 0);
    if (lane == 0) {
        if (any_l) atomicOr(&flags->any_live, 1u);
        if (any_c) atomicOr(&flags->any_conflict, 1u);
        if (any_u) atomicOr(&flags->any_unsettled, 1u);
    }
}

__global__ void __launch_bounds__(256) k_status_step(uint64_t *changed, uint64_t *conf_step, uint64_t *live,
                                                     uint64_t *conflict, uint64_t *unsettled, RunFlags *flags,
                                                     uint32_t words, uint64_t n_lanes, int count_conflicts, int is_last,
                                                     RunFlags *next_flags, const uint32_t *counts) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w == 0) {  // the slot of the next application (its previous content was read eight applications ago)
        next_flags->any_live = 0;
        next_flags->any_conflict = 0;
        next_flags->any_unsettled = 0;
        flags->active = counts ? counts[0] + counts[1] + counts[2] + counts[3] : 0xFFFFFFFFu;
    }
    status_step_body(changed, conf_step, live, conflict, unsettled, flags, words, n_lanes, count_conflicts, is_last, w, threadIdx.x & 31);
}

__global__ void __launch_bounds__(256) k_init_status(uint64_t *changed, uint64_t *conf_step, uint64_t *live, uint64_t *conflict,
                                                     uint64_t *unsettled, RunFlags *flags, uint32_t words, uint64_t n_lanes,
                                                     int all_conflict) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < 8) {  // all flag slots of the unit-delay loop; slot 0 doubles as the result of a run that never steps
        flags[w].any_live = 0;
        flags[w].any_conflict = (all_conflict && w == 0) ? 1u : 0u;
        flags[w].any_unsettled = 0;
    }
    if (w >= words) return;
    const uint64_t mask = lane_mask_of(w, n_lanes);
    changed[w] = 0;
    conf_step[w] = 0;
    live[w] = mask;
    conflict[w] = all_conflict ? mask : 0;
    unsettled[w] = 0;
}

__global__ void k_clear_flags(RunFlags *flags) {
    flags->any_live = 0;
    flags->any_conflict = 0;
    flags->any_unsettled = 0;
}

// Levelised runs: conflicts can only come from drives on gate-driven rows.
__global__ void __launch_bounds__(256) k_levelised_status(uint64_t *conflict, uint64_t *unsettled, RunFlags *flags,
                                                          uint32_t words, uint64_t n_lanes) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t cf = 0;
    if (w < words) {
        cf = conflict[w] & lane_mask_of(w, n_lanes);
        conflict[w] = cf;
        unsettled[w] = 0;
    }
    const unsigned any_c = __ballot_sync(0xFFFFFFFFu, cf != 0);
    if ((threadIdx.x & 31) == 0 && any_c) atomicOr(&flags->any_conflict, 1u);
}

// set_wire_drive, batched: every pending (bit-net, value bit, valid bit) is broadcast to all lanes of
// its drive row.  One launch per flush, however many wires were driven since the last run.
struct PendingBit {
    uint64_t *val, *vld;  // drive row of the bit-net
    uint32_t flags;       // bit 0 = plane 0 (value), bit 1 = plane 1 (valid)
    uint32_t pad;
};

__global__ void __launch_bounds__(256) k_broadcast_bits(const PendingBit *bits, uint32_t n_bits, uint32_t stride,
                                                        uint32_t word_blocks) {
    const uint32_t j = blockIdx.x / word_blocks, w = (blockIdx.x % word_blocks) * blockDim.x + threadIdx.x;
    if (w >= stride || j >= n_bits) return;
    const PendingBit b = bits[j];
    b.val[w] = (b.flags & 1) ? ~0ull : 0ull;
    b.vld[w] = (b.flags & 2) ? ~0ull : 0ull;
}

__global__ void __launch_bounds__(256) k_fill_stimulus(uint64_t *const *row_val, uint64_t *const *row_vld, uint32_t n,
                                                       uint32_t words, uint64_t seed, uint64_t word_base, int mode) {
    // one 128-bit pair of lane-words per thread (rows are 16-byte aligned and hold an even number of words)
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const uint32_t i = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= words || i >= n) return;
    const uint64_t k = seed ^ ((uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull);
    ulonglong2 v, m;
    v.x = splitmix64(k ^ ((word_base + w) * 0xD1B54A32D192ED03ull));
    v.y = splitmix64(k ^ ((word_base + w + 1) * 0xD1B54A32D192ED03ull));
    m.x = mode == 0 ? ~0ull : (splitmix64(v.x ^ 1) | splitmix64(v.x ^ 2));
    m.y = mode == 0 ? ~0ull : (splitmix64(v.y ^ 1) | splitmix64(v.y ^ 2));
    st_pair(row_val[i] + w, v);
    st_pair(row_vld[i] + w, m);
}

// rows[] of (val, vld) -> dense planes [n][dst_stride].  VEC: one 128-bit pair of lane-words per thread (every offset
// even, every base 16-byte aligned — checked by the caller); the destination may be peer-GPU memory (the fused output gather).
template <bool VEC>
__global__ void __launch_bounds__(256) k_gather_rows(const uint32_t *rows, uint32_t n, const uint64_t *val, const uint64_t *vld,
                                                     uint64_t *dst_val, uint64_t *dst_vld, uint32_t stride, uint32_t word0,
                                                     uint32_t n_words, size_t dst_stride, const uint8_t *kv,
                                                     const uint32_t *all_valid) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 2 : 1);
    const uint32_t i = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= n_words || i >= n) return;
    const size_t s = (size_t)rows[i] * stride + word0 + w, o = (size_t)i * dst_stride + w;
    const bool elided = kv && (*all_valid || kv[rows[i]]);
    if (VEC) {
        st_pair(dst_val + o, ld_stream(val + s));
        st_pair(dst_vld + o, elided ? make_ulonglong2(~0ull, ~0ull) : ld_stream(vld + s));
    } else {
        dst_val[o] = val[s];
        dst_vld[o] = elided ? ~0ull : vld[s];
    }
}

// the same rows to up to kMaxGatherDst destinations: read once, stored to every destination (own planes, peer GPUs' planes)
static const uint32_t kMaxGatherDst = 9;
struct GatherDsts {
    uint64_t *val[kMaxGatherDst], *vld[kMaxGatherDst];
    uint64_t stride[kMaxGatherDst];
    uint32_t n;
};
template <bool VEC>
__global__ void __launch_bounds__(256) k_gather_rows_multi(const uint32_t *rows, uint32_t n, const uint64_t *val, const uint64_t *vld,
                                                           uint32_t stride, uint32_t n_words, const GatherDsts d, const uint8_t *kv,
                                                           const uint32_t *all_valid) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 2 : 1);
    const uint32_t i = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= n_words || i >= n) return;
    const size_t s = (size_t)rows[i] * stride + w;
    const bool elided = kv && (*all_valid || kv[rows[i]]);
    if (VEC) {
        const ulonglong2 v = ld_stream(val + s), m = elided ? make_ulonglong2(~0ull, ~0ull) : ld_stream(vld + s);
        for (uint32_t k = 0; k < d.n; k++) {
            const size_t o = (size_t)i * d.stride[k] + w;
            st_pair(d.val[k] + o, v);
            st_pair(d.vld[k] + o, m);
        }
    } else {
        const uint64_t v = val[s], m = elided ? ~0ull : vld[s];
        for (uint32_t k = 0; k < d.n; k++) {
            const size_t o = (size_t)i * d.stride[k] + w;
            d.val[k][o] = v;
            d.vld[k][o] = m;
        }
    }
}

// dense planes [n][src_stride] -> drive rows given by pointer tables
template <bool VEC>
__global__ void __launch_bounds__(256) k_scatter_rows(uint64_t *const *row_val, uint64_t *const *row_vld, uint32_t n,
                                                      const uint64_t *src_val, const uint64_t *src_vld, uint32_t words,
                                                      size_t src_stride) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 2 : 1);
    const uint32_t i = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= words || i >= n) return;
    const size_t s = (size_t)i * src_stride + w;
    if (VEC) {
        st_pair(row_val[i] + w, ld_stream(src_val + s));
        st_pair(row_vld[i] + w, ld_stream(src_vld + s));
    } else {
        row_val[i][w] = src_val[s];
        row_vld[i][w] = src_vld[s];
    }
}

// SimState bit-packing of lane 0 (reference crates/digilogic_netcode/src/lib.rs:241-270): bit k
// of the packed planes = bit 0 / lane 0 of rows[k].  One thread builds one output byte.
__global__ void __launch_bounds__(256) k_pack_lane0(const uint32_t *rows, uint32_t n_bits, const uint64_t *val,
                                                    const uint64_t *vld, uint32_t stride, uint8_t *out0, uint8_t *out1,
                                                    uint32_t n_bytes) {
    const uint32_t byte = blockIdx.x * blockDim.x + threadIdx.x;
    if (byte >= n_bytes) return;
    uint32_t b0 = 0, b1 = 0;
    for (uint32_t j = 0; j < 8; j++) {
        const uint32_t k = byte * 8 + j;
        if (k < n_bits) {
            const size_t s = (size_t)(rows ? rows[k] : k) * stride;
            b0 |= (uint32_t)(val[s] & 1) << j;
            b1 |= (uint32_t)(vld[s] & 1) << j;
        }
    }
    out0[byte] = (uint8_t)b0;
    out1[byte] = (uint8_t)b1;
}

// ---------------------------------------------------------------------------- device-side run_sim loop
//
// The whole unit-delay loop of one run_sim as ONE cooperative launch: per application the same phases as the
// host-driven loop (drives -> frontier scan | listed gates, refresh copies, wide gates | tri-state resolve |
// lane bookkeeping), separated by grid barriers instead of kernel boundaries, and the stop decision
// ("no lane is live any more", or the step budget) is taken on the device.  An application in which little
// happens — an oscillating latch in a few lanes, the tail of a settle — costs three grid barriers, not a dozen
// launches and a host round trip.
struct LoopResult {
    uint64_t k;  // index of the last application performed
    int32_t cur, chg_w;
    uint32_t any_conflict, any_unsettled;
    uint32_t bailed;  // 1: stopped before application k+1 because its work list is too long for this kernel; the host loop goes on
    uint32_t pad;
    uint64_t skipped;  // applications fast-forwarded over (exact period-2 oscillation), included in k
};

struct LoopArgs {
    const uint32_t *fan0, *fan1, *fan2;
    const uint8_t *kind2;
    WideArgs wide;     // netlist part; the plane pointers are set per application
    ResolveArgs res;   // likewise
    uint64_t *val[2], *vld[2];
    const uint64_t *drv_val, *drv_vld;
    uint8_t *chg[2];
    uint4 *list2, *list3;
    uint32_t *listc, *listw;
    uint32_t *cnt;  // list counters, [2][4]: application parity x {one/two-input, three-input, refresh, -}; zero on entry
    uint64_t *changed, *conf_step, *live, *conflict, *unsettled;
    RunFlags *flags;  // kFlagSlots slots, all clear on entry
    LoopResult *result;
    uint32_t n_src, n_g2, n_fast, n_wide, n_res, stride, pairs, pair_blocks, rowchg_bytes;
    uint64_t n_lanes, k0, last;
    uint64_t bail_row_words;  // hand back to the host loop when listed gates x row words exceeds this
    int32_t cur, chg_w, frontier_valid;
    uint32_t sem;  // SemFlags of the netlist
    uint32_t *p2;  // period-2 detector, kFlagSlots words (slot = application % kFlagSlots), zero on entry; nullptr = off
};

__global__ void __launch_bounds__(256, 2) k_unit_delay_loop(const LoopArgs a) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t bid = blockIdx.x, nblk = gridDim.x;
    const uint32_t pb = bid % a.pair_blocks, gbi = bid / a.pair_blocks, gbs = nblk / a.pair_blocks;
    const uint32_t lt = threadIdx.y * blockDim.x + threadIdx.x;  // linear thread id: warps are 32 consecutive ones
    uint64_t k = a.k0, skipped = 0;
    int cur = a.cur, cw = a.chg_w;
    bool bailed = false;
    const uint64_t first_k = k + 1;
    // The work lists of application k are built at the end of application k-1 (their counters: a.cnt + 4 * (k & 1)),
    // so an application is two grid barriers: [drives, listed gates, refresh copies, wide gates] | [tri-state resolve |]
    // [lane bookkeeping, flag clean-up, scan for the next application].  Prologue: the scan for the first one.
    for (uint32_t base = bid * 256; base < a.n_fast; base += nblk * 256)
        frontier_scan_body(a.fan0, a.fan1, a.fan2, a.n_g2, a.n_fast, a.n_src, a.kind2, a.frontier_valid ? a.chg[cw ^ 1] : nullptr,
                           a.list2, a.list3, a.listc, a.cnt + 4 * (first_k & 1), lt, base + lt);
    for (uint32_t base = bid * 256; base < a.n_wide; base += nblk * 256)
        frontier_scan_wide_body(a.wide.off, a.wide.in, a.n_wide, a.wide.row_base, a.frontier_valid ? a.chg[cw ^ 1] : nullptr, a.listw,
                                a.cnt + 4 * (first_k & 1) + 3, lt, base + lt);
    grid.sync();
    for (;;) {
        k++;
        const bool is_last = k == a.last;
        const uint64_t *AV = a.val[cur], *AM = a.vld[cur];
        uint64_t *BV = a.val[cur ^ 1], *BM = a.vld[cur ^ 1];
        uint8_t *chg_cur = a.chg[cw];
        const uint32_t *counts = a.cnt + 4 * (k & 1);
        RunFlags *flags = a.flags + (k % 8);
        uint32_t *p2 = a.p2 ? a.p2 + (k % 8) : nullptr;
        if (k > first_k) {  // a long work list is better served by the separate, fully occupied kernels
            const uint64_t listed = (uint64_t)counts[0] + counts[1] + counts[2] + counts[3];
            if (listed * a.stride > a.bail_row_words) {
                bailed = true;
                k--;
                break;
            }
        }

        // ---- phase A: drives (they cannot change inside a run: two applications fill both buffers), the listed
        //      gates, the rows to refresh, the wide gates
        if (a.n_src && k < first_k + 2)
            for (uint32_t rb = gbi; rb * blockDim.y < a.n_src; rb += gbs)
                apply_drive_body<true>(a.drv_val, a.drv_vld, AV, AM, BV, BM, a.changed, a.n_src, a.stride, a.pairs, nullptr, 0, chg_cur,
                                       pb, rb);
        if (bid == 0 && lt < 4) a.cnt[4 * ((k + 1) & 1) + lt] = 0;  // counters of the next application (last read an application ago)
        {
            Eval2Args e{a.fan0, a.fan1, a.fan2, a.kind2, AV, AM, BV, BM, a.changed, 0, a.n_fast, a.n_src, a.stride, a.pairs,
                        a.pair_blocks, nullptr, nullptr, chg_cur, nullptr};
            e.sem = a.sem;
            e.p2diff = p2;
            if (a.n_g2) eval_list_body<2>(e, a.list2, counts, bid, nblk);
            if (a.n_fast > a.n_g2) eval_list_body<3>(e, a.list3, counts + 1, bid, nblk);
            copy_list_body(a.listc, counts + 2, AV, AM, BV, BM, a.stride, a.pairs, a.pair_blocks, bid, nblk, p2);
            if (a.n_wide) {
                WideArgs w = a.wide;
                w.in_val = AV; w.in_vld = AM; w.out_val = BV; w.out_vld = BM;
                w.chg_prev = nullptr; w.chg_cur = chg_cur;
                w.p2diff = p2;
                eval_wide_list_body(w, a.listw, counts + 3, bid, nblk);
            }
        }
        grid.sync();
        if (a.n_res) {  // tri-state nets see their drivers' rows of this application
            ResolveArgs r = a.res;
            r.a_val = AV; r.a_vld = AM; r.b_val = BV; r.b_vld = BM;
            r.chg_cur = chg_cur;
            r.p2diff = p2;
            for (uint32_t rb = gbi; rb * blockDim.y < a.n_res; rb += gbs) resolve_body(r, pb, rb);
            grid.sync();
        }

        // ---- phase B: lane bookkeeping; the flags of the previous application become the clean buffer of the next;
        //      the work lists of the next application from the flags of this one
        for (uint32_t base = bid * 256; base < a.stride; base += nblk * 256)
            status_step_body(a.changed, a.conf_step, a.live, a.conflict, a.unsettled, flags, a.stride, a.n_lanes, is_last ? 0 : 1,
                             is_last ? 1 : 0, base + lt, lt & 31);
        {
            uint4 *z = (uint4 *)a.chg[cw ^ 1];
            for (uint32_t i = bid * 256 + lt; i < a.rowchg_bytes / 16; i += nblk * 256) z[i] = make_uint4(0, 0, 0, 0);
        }
        if (bid == 0 && lt == 0) {
            RunFlags *nf = a.flags + ((k + 1) % 8);
            nf->any_live = 0;
            nf->any_conflict = 0;
            nf->any_unsettled = 0;
            if (a.p2) a.p2[(k + 1) % 8] = 0;
        }
        if (!is_last)
            for (uint32_t base = bid * 256; base < a.n_fast; base += nblk * 256)
                frontier_scan_body(a.fan0, a.fan1, a.fan2, a.n_g2, a.n_fast, a.n_src, a.kind2, chg_cur, a.list2, a.list3, a.listc,
                                   a.cnt + 4 * ((k + 1) & 1), lt, base + lt);
        if (!is_last)
            for (uint32_t base = bid * 256; base < a.n_wide; base += nblk * 256)
                frontier_scan_wide_body(a.wide.off, a.wide.in, a.n_wide, a.wide.row_base, chg_cur, a.listw, a.cnt + 4 * ((k + 1) & 1) + 3,
                                        lt, base + lt);
        grid.sync();
        const uint32_t any_live = *(volatile uint32_t *)&flags->any_live;
        cw ^= 1;
        if (is_last) break;  // check-only application: the state stays W_{N+1}
        cur ^= 1;
        if (!any_live) break;
        // Exact period-2 oscillation: every row this application wrote equals what its buffer held, i.e. W_k = W_{k-2} on all
        // rows (rows not listed are equal by the frontier invariant) and the drives have been in both buffers since the
        // second application of this launch.  G is a function of the state alone, so W_{k+2j} = W_k and W_{k+2j+1} = W_{k-1}
        // for all j, with the conflicts, live lanes, changed-row flags and work lists of applications k-1 and k repeating:
        // jumping ahead by an even number of applications is the identity on everything this loop keeps.  Only the flag slots
        // indexed by the application number have to be brought along.
        if (p2 && k >= first_k + 2 && *(volatile uint32_t *)p2 == 0) {
            const uint64_t jump = ((a.last - 1 - k) / 2) * 2;
            if (jump >= 8) {
                const uint64_t k2 = k + jump;
                if (bid == 0 && lt == 0) {
                    RunFlags *dst = a.flags + (k2 % 8), *nf = a.flags + ((k2 + 1) % 8);
                    *dst = *flags;  // (conflict flags are cumulative: the final read finds them in the slot of the last application)
                    nf->any_live = 0;
                    nf->any_conflict = 0;
                    nf->any_unsettled = 0;
                    a.p2[k2 % 8] = 0;
                    a.p2[(k2 + 1) % 8] = 0;
                }
                skipped += jump;
                k = k2;
                grid.sync();
            }
        }
    }
    if (bid == 0 && lt == 0) {
        const RunFlags *flags = a.flags + (k % 8);
        a.result->k = k;
        a.result->cur = cur;
        a.result->chg_w = cw;
        a.result->any_conflict = *(volatile uint32_t *)&flags->any_conflict;
        a.result->any_unsettled = *(volatile uint32_t *)&flags->any_unsettled;
        a.result->bailed = bailed ? 1u : 0u;
        a.result->skipped = skipped;
    }
}

// ---------------------------------------------------------------------------- resident kernel
//
// Small netlists: the whole wire-state store of a few lane-words fits in shared memory, so one
// CTA runs the complete run_sim loop for its lane-words on chip — every application of G reads
// its fan-in from shared memory, __syncthreads() separates applications, and nothing but the
// drives (in) and the final state + lane status (out) touches HBM.  Lanes are independent, so
// CTAs never talk to each other and each stops at its own fixed point.  This is the path of the
// reference's real use (GUI: one lane, max_steps = 10 000, latches that may oscillate).
static constexpr uint32_t kInlineBits = 96;
struct ResidentArgs {
    const uint32_t *fan0, *fan1, *fan2;
    const uint8_t *kind2;
    const uint32_t *woff, *win;
    const uint8_t *wkind, *wnsel, *waux;
    const uint32_t *res_off, *res_in; // tri-state nets: driver rows per resolved row (see k_resolve)
    const int32_t *res_xslot;
    uint32_t n_res;
    const uint32_t *xrows;            // rows with an external drive on a gate output
    const uint64_t *xval, *xvld;      // their drives [n_extra][stride]
    const uint64_t *dval, *dvld;      // source drives [n_src][stride]
    uint64_t *val, *vld;              // store (current buffer), read at start, written at the end
    uint64_t *conflict, *unsettled;   // lane status words [stride]
    RunFlags *flags;
    unsigned long long *max_apps;     // max applications over CTAs
    uint32_t n_src, n_g2, n_g3, n_wide, n_extra, n_rows;   // n_g2: one/two-input gates, n_g3: three-input gates
    uint32_t stride, wpc;             // words per row; lane-words per CTA
    uint64_t n_lanes;
    uint64_t last;                    // index of the final, check-only application
    int cold;
    // Acyclic single-driver netlist and a step budget that covers its depth: the fixed point is computed
    // once, level by level, in place (the same rule as Engine::run_levelised), with a CTA barrier per
    // level instead of a launch per level.
    const Level *levels;
    uint32_t n_levels;
    int levelised;
    int lenient;  // A.6-2 alternative: agreeing valid drivers do not conflict
    // small runs (the one-lane server path) also leave the bit-packed lane-0 snapshot get_wire_state reads (CTA 0 packs it)
    uint8_t *snap0, *snap1;
    uint32_t snap_bytes;
    uint32_t sem;  // SemFlags of the netlist
    int period2;   // B2_OPT_PERIOD2_SKIP
    // One-CTA runs (the reference server's one-lane cycle: set a few drives, eval, read every net): the drives queued by
    // set_wire_drive travel in the kernel's arguments and are broadcast into their drive rows by the kernel itself, and the
    // run's flags (like the snapshot above) are stored straight into page-locked host memory — one launch and one
    // synchronisation per eval, no copy in either direction.
    RunFlags *host_flags;  // [2] on the host, or nullptr: flags through `flags` / `max_apps` and a device-to-host copy
    uint32_t n_inline;
    PendingBit inl[kInlineBits];
};

__device__ __forceinline__ void wide_gate_word(const ResidentArgs &a, uint32_t g, const uint64_t *av, const uint64_t *am, uint32_t wpc,
                                               uint32_t j, u64 &rv, u64 &rm) {
    const uint32_t b = a.woff[g], e = a.woff[g + 1], kind = a.wkind[g];
    if (kind == K_ADD || kind == K_SUB) {
        const uint32_t w = (e - b) / 2, bit = a.wnsel[g];
        const bool sub = kind == K_SUB;
        u64 ok = ~0ull, carry = sub ? ~0ull : 0ull, sum = 0;
        for (uint32_t t = 0; t < w; t++) {
            const size_t ra = (size_t)a.win[b + t] * wpc + j, rb = (size_t)a.win[b + w + t] * wpc + j;
            ok &= am[ra] & am[rb];
            if (t <= bit) {
                const u64 x = av[ra], y = sub ? ~av[rb] : av[rb];
                if (t == bit) sum = x ^ y ^ carry;
                else carry = (x & y) | (carry & (x ^ y));
            }
        }
        rm = ok;
        rv = sum | ~ok;
    } else if (kind >= K_MUL && kind <= K_ASHR) {
        auto v = [&](uint32_t i) { return (u64)av[(size_t)a.win[b + i] * wpc + j]; };
        auto m = [&](uint32_t i) { return (u64)am[(size_t)a.win[b + i] * wpc + j]; };
        arith_word(kind, e - b, a.wnsel[g], a.waux[g], v, m, rv, rm);
    } else if (kind == K_REG) {
        const size_t rd = (size_t)a.win[b] * wpc + j, re = (size_t)a.win[b + 1] * wpc + j, rc = (size_t)a.win[b + 2] * wpc + j,
                     rp = (size_t)a.win[b + 3] * wpc + j, ro = (size_t)(a.n_src + a.n_g2 + a.n_g3 + g) * wpc + j;
        reg_word(av[rd], am[rd], av[re], am[re], av[rc], am[rc], av[rp], am[rp], av[ro], am[ro], rv, rm);
    } else if (kind == K_MUX) {
        const uint32_t ns = a.wnsel[g];
        u64 allv = ~0ull, pv = 0, pm = 0;
        for (uint32_t s = 0; s < ns; s++) allv &= am[(size_t)a.win[b + s] * wpc + j];
        for (uint32_t d = 0; b + ns + d < e; d++) {
            u64 match = ~0ull;
            for (uint32_t s = 0; s < ns; s++) match &= ~(av[(size_t)a.win[b + s] * wpc + j] ^ (((d >> s) & 1) ? ~0ull : 0ull));
            const size_t r = (size_t)a.win[b + ns + d] * wpc + j;
            pv |= match & av[r];
            pm |= match & am[r];
        }
        rm = allv & pm;
        rv = (allv & pv) | ((a.sem & SEM_MUX_Z) ? ~allv : ~rm);  // A.6-4: High-Z on the selected input reads X, or passes through
    } else {
        const uint32_t base = kind >= K_NAND ? kind - 3 : kind;
        size_t r = (size_t)a.win[b] * wpc + j;
        rv = av[r];
        rm = am[r];
        for (uint32_t i = b + 1; i < e; i++) {
            r = (size_t)a.win[i] * wpc + j;
            gate_word(base, rv, rm, av[r], am[r], rv, rm, a.sem);
        }
        if (kind >= K_NAND) rv = ~rv | ~rm;
    }
}

__global__ void __launch_bounds__(1024, 1) k_resident_run(const ResidentArgs a) {
    extern __shared__ uint64_t smem[];
    const uint32_t wpc = a.wpc, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t w0 = blockIdx.x * wpc;
    const size_t plane = (size_t)a.n_rows * wpc;
    uint64_t *AV = smem, *AM = smem + plane, *BV = smem + 2 * plane, *BM = smem + 3 * plane;
    uint64_t *s_changed = smem + 4 * plane, *s_conf = s_changed + wpc, *s_live = s_conf + wpc, *s_conflict = s_live + wpc,
             *s_unsettled = s_conflict + wpc;
    __shared__ unsigned s_any, s_p2[2];
    if (threadIdx.x < 2) s_p2[threadIdx.x] = 0;
    if (a.n_inline) {  // one CTA (the host sees to it): queued drives into their rows, like k_broadcast_bits
        for (uint32_t i = tid; i < a.n_inline * a.stride; i += nt) {
            const PendingBit &b = a.inl[i / a.stride];
            b.val[i % a.stride] = (b.flags & 1) ? ~0ull : 0ull;
            b.vld[i % a.stride] = (b.flags & 2) ? ~0ull : 0ull;
        }
        __threadfence_block();
        __syncthreads();
    }

    // ---- load: the current state, or the cold start W_1 = resolve(D, all outputs High-Z) = D
    //      (a levelised pass recomputes every row: only the source rows are loaded, from their drives)
    for (size_t i = tid; i < (a.levelised ? (size_t)a.n_src * wpc : plane); i += nt) {
        const uint32_t row = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
        u64 v = 0, m = 0;
        if (w < a.stride) {
            const size_t o = (size_t)row * a.stride + w;
            if (!a.cold && !a.levelised) {
                v = a.val[o];
                m = a.vld[o];
            } else if (row < a.n_src) {
                v = a.dval[o];
                m = a.dvld[o];
            } else if (a.sem & SEM_INIT_X) {
                v = ~0ull;  // A.6-5 alternative: component outputs start Undefined
            }
        }
        AV[i] = v;
        AM[i] = m;
    }
    if (tid < wpc) {
        s_changed[tid] = 0;
        s_conf[tid] = 0;
        s_live[tid] = (w0 + tid < a.stride) ? lane_mask_of(w0 + tid, a.n_lanes) : 0;
        s_conflict[tid] = 0;
        s_unsettled[tid] = 0;
    }
    __syncthreads();
    const uint32_t res_base = a.n_src + a.n_g2 + a.n_g3 + a.n_wide;
    if (a.levelised) {
        for (uint32_t l = 0; l < a.n_levels; l++) {
            const Level lv = a.levels[l];
            const uint32_t n2 = lv.g2_end - lv.g2_begin, n3 = lv.g3_end - lv.g3_begin, nw = lv.wide_end - lv.wide_begin;
            const size_t total = (size_t)(n2 + n3 + nw) * wpc;
            for (size_t i = tid; i < total; i += nt) {
                const uint32_t gi = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc);
                u64 rv, rm;
                uint32_t row;
                if (gi < n2) {
                    const uint32_t g = lv.g2_begin + gi;
                    const size_t r0 = (size_t)a.fan0[g] * wpc + j, r1 = (size_t)a.fan1[g] * wpc + j;
                    gate_word(a.kind2[g], AV[r0], AM[r0], AV[r1], AM[r1], rv, rm, a.sem);
                    row = a.n_src + g;
                } else if (gi < n2 + n3) {
                    const uint32_t g = lv.g3_begin + (gi - n2), kind = a.kind2[g], base = kind >= K_NAND ? kind - 3 : kind;
                    const size_t r0 = (size_t)a.fan0[g] * wpc + j, r1 = (size_t)a.fan1[g] * wpc + j, r2 = (size_t)a.fan2[g] * wpc + j;
                    gate_word(base, AV[r0], AM[r0], AV[r1], AM[r1], rv, rm, a.sem);
                    gate_word(base, rv, rm, AV[r2], AM[r2], rv, rm, a.sem);
                    if (kind >= K_NAND) rv = ~rv | ~rm;
                    row = a.n_src + g;
                } else {
                    const uint32_t g = lv.wide_begin + (gi - n2 - n3);
                    wide_gate_word(a, g, AV, AM, wpc, j, rv, rm);
                    row = a.n_src + a.n_g2 + a.n_g3 + g;
                }
                AV[(size_t)row * wpc + j] = rv;  // in place: gates of one level never read each other
                AM[(size_t)row * wpc + j] = rm;
            }
            __syncthreads();
        }
        // drives on gate outputs: a gate output is never High-Z, so the lane conflicts where the drive is not High-Z
        for (size_t i = tid; i < (size_t)a.n_extra * wpc; i += nt) {
            const uint32_t x = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
            if (w < a.stride && a.xrows[x] != 0xFFFFFFFFu) {
                const size_t r = (size_t)a.xrows[x] * wpc + j;
                const u64 dv = a.xval[(size_t)x * a.stride + w], dm = a.xvld[(size_t)x * a.stride + w];
                const u64 c = conflict_of(dv, dm, AV[r], AM[r], a.lenient);
                AV[r] |= dv;
                AM[r] |= dm;
                if (c) atomicOr((unsigned long long *)&s_conflict[j], c);
            }
        }
        __syncthreads();
        if (tid < wpc) s_conflict[tid] &= s_live[tid];  // lanes in use
    } else if (a.cold) {
        const bool init_x = (a.sem & SEM_INIT_X) != 0;  // outputs start Undefined: they take part in W_1 and can conflict there
        for (size_t i = tid; i < (size_t)a.n_extra * wpc; i += nt) {
            const uint32_t x = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
            if (w < a.stride && a.xrows[x] != 0xFFFFFFFFu) {
                const size_t r = (size_t)a.xrows[x] * wpc + j;
                const u64 dv = a.xval[(size_t)x * a.stride + w], dm = a.xvld[(size_t)x * a.stride + w];
                if (init_x) {
                    const u64 c = conflict_of(dv, dm, AV[r], AM[r], a.lenient);
                    AV[r] |= dv;
                    AM[r] |= dm;
                    if (c) atomicOr((unsigned long long *)&s_conf[j], c);
                } else {
                    AV[r] = dv;
                    AM[r] = dm;
                }
            }
        }
        for (size_t i = tid; i < (size_t)a.n_res * wpc; i += nt) {
            const uint32_t r = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
            const int32_t slot = a.res_xslot[r];
            u64 v = 0, m = 0, c = 0;
            if (w < a.stride && slot >= 0) {
                v = a.xval[(size_t)slot * a.stride + w];
                m = a.xvld[(size_t)slot * a.stride + w];
            }
            if (init_x)
                for (uint32_t q = a.res_off[r]; q < a.res_off[r + 1]; q++) {  // every driver contributes an Undefined
                    c |= conflict_of(v, m, ~0ull, 0, a.lenient);
                    v = ~0ull;
                }
            if (w < a.stride && (slot >= 0 || init_x)) {
                AV[(size_t)(res_base + r) * wpc + j] = v;
                AM[(size_t)(res_base + r) * wpc + j] = m;
                if (c) atomicOr((unsigned long long *)&s_conf[j], c);
            }
        }
        __syncthreads();
    }

    uint64_t k = a.cold ? 1 : 0;
    unsigned long long apps = 0;
    const size_t n_items = (size_t)(a.n_src + a.n_g2 + a.n_g3 + a.n_wide) * wpc;
    // (a cold start without components is already final: W_1 = D; a warm one still takes the new drives)
    if (a.levelised) {
        apps = 1;
    } else if (a.n_g2 + a.n_g3 + a.n_wide > 0 || !a.cold) {
        for (;;) {
            k++;
            apps++;
            const bool is_last = k == a.last;
            u64 d2 = 0;  // does this application write anything that differs from the state two applications ago?
            // ---- B = G(A)
            for (size_t i = tid; i < n_items; i += nt) {
                const uint32_t row = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
                u64 rv, rm;
                bool counts = true;
                if (row < a.n_src) {
                    rv = rm = 0;
                    if (w < a.stride) {
                        rv = a.dval[(size_t)row * a.stride + w];
                        rm = a.dvld[(size_t)row * a.stride + w];
                    }
                } else if (row < a.n_src + a.n_g2) {
                    const uint32_t g = row - a.n_src;
                    const size_t r0 = (size_t)a.fan0[g] * wpc + j, r1 = (size_t)a.fan1[g] * wpc + j;
                    gate_word(a.kind2[g], AV[r0], AM[r0], AV[r1], AM[r1], rv, rm, a.sem);
                    counts = a.kind2[g] != K_RAW;  // a register's clock history is internal state: it keeps no lane alive
                } else if (row < a.n_src + a.n_g2 + a.n_g3) {
                    const uint32_t g = row - a.n_src, kind = a.kind2[g], base = kind >= K_NAND ? kind - 3 : kind;
                    const size_t r0 = (size_t)a.fan0[g] * wpc + j, r1 = (size_t)a.fan1[g] * wpc + j, r2 = (size_t)a.fan2[g] * wpc + j;
                    gate_word(base, AV[r0], AM[r0], AV[r1], AM[r1], rv, rm, a.sem);
                    gate_word(base, rv, rm, AV[r2], AM[r2], rv, rm, a.sem);
                    if (kind >= K_NAND) rv = ~rv | ~rm;
                } else {
                    wide_gate_word(a, row - a.n_src - a.n_g2 - a.n_g3, AV, AM, wpc, j, rv, rm);
                }
                const u64 d = (rv ^ AV[i]) | (rm ^ AM[i]);
                d2 |= (rv ^ BV[i]) | (rm ^ BM[i]);
                BV[i] = rv;
                BM[i] = rm;
                if (d && counts) atomicOr((unsigned long long *)&s_changed[j], d);
            }
            if (d2) s_p2[k & 1] = 1;
            __syncthreads();
            // ---- tri-state nets: wire = resolve(drive, driver rows of B), conflict where two are not High-Z
            if (a.n_res) {
                for (size_t i = tid; i < (size_t)a.n_res * wpc; i += nt) {
                    const uint32_t r = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
                    u64 v = 0, m = 0, c = 0;
                    const int32_t slot = a.res_xslot[r];
                    if (slot >= 0 && w < a.stride) {
                        v = a.xval[(size_t)slot * a.stride + w];
                        m = a.xvld[(size_t)slot * a.stride + w];
                    }
                    for (uint32_t q = a.res_off[r]; q < a.res_off[r + 1]; q++) {
                        const size_t o = (size_t)a.res_in[q] * wpc + j;
                        c |= conflict_of(v, m, BV[o], BM[o], a.lenient);
                        v |= BV[o];
                        m |= BM[o];
                    }
                    const size_t idx = (size_t)(res_base + r) * wpc + j;
                    const u64 d = (v ^ AV[idx]) | (m ^ AM[idx]);
                    BV[idx] = v;
                    BM[idx] = m;
                    if (d) atomicOr((unsigned long long *)&s_changed[j], d);
                    if (c) atomicOr((unsigned long long *)&s_conf[j], c);
                }
                __syncthreads();
            }
            // ---- drives on gate outputs: merge + conflict
            if (a.n_extra) {
                for (size_t i = tid; i < (size_t)a.n_extra * wpc; i += nt) {
                    const uint32_t x = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
                    if (w < a.stride && a.xrows[x] != 0xFFFFFFFFu) {
                        const size_t r = (size_t)a.xrows[x] * wpc + j;
                        const u64 dv = a.xval[(size_t)x * a.stride + w], dm = a.xvld[(size_t)x * a.stride + w];
                        const u64 c = conflict_of(dv, dm, BV[r], BM[r], a.lenient);
                        BV[r] |= dv;
                        BM[r] |= dm;
                        if (c) atomicOr((unsigned long long *)&s_conf[j], c);
                    }
                }
                __syncthreads();
            }
            // ---- per-lane bookkeeping (same rules as k_status_step)
            if (tid == 0) {
                s_any = 0;
                s_p2[(k + 1) & 1] = 0;
            }
            __syncthreads();
            if (tid < wpc) {
                u64 cf = s_conflict[tid], lv = s_live[tid];
                if (!is_last) cf |= s_conf[tid] & lv;
                lv &= s_changed[tid] & ~cf;
                s_live[tid] = lv;
                s_conflict[tid] = cf;
                if (is_last) s_unsettled[tid] = lv;
                s_changed[tid] = 0;
                s_conf[tid] = 0;
                if (lv) atomicOr(&s_any, 1u);
            }
            __syncthreads();
            if (is_last) break;  // check-only application: the state stays W_{N+1}
            uint64_t *t;
            t = AV; AV = BV; BV = t;
            t = AM; AM = BM; BM = t;
            if (!s_any) break;
            // Exact period-2 oscillation (W_k = W_{k-2} on every gate and source row; tri-state rows follow from them): the
            // step function depends on the state alone, so jumping ahead by an even number of applications changes nothing
            // this loop keeps — see k_unit_delay_loop.  apps >= 3: the buffer compared with was written by this launch.
            if (a.period2 && apps >= 3 && a.n_extra == 0 && !s_p2[k & 1]) {
                const uint64_t jump = ((a.last - 1 - k) / 2) * 2;
                if (jump >= 8) k += jump;
            }
        }
    }
    // ---- write back
    for (size_t i = tid; i < plane; i += nt) {
        const uint32_t row = (uint32_t)(i / wpc), j = (uint32_t)(i % wpc), w = w0 + j;
        if (w < a.stride) {
            a.val[(size_t)row * a.stride + w] = AV[i];
            a.vld[(size_t)row * a.stride + w] = AM[i];
        }
    }
    if (a.snap0 && a.host_flags) {
        // straight into page-locked host memory: 64 rows per store (a byte store each would be a PCIe write each)
        for (uint32_t q = tid; q < (a.snap_bytes + 7) / 8; q += nt) {
            uint64_t b0 = 0, b1 = 0;
            for (uint32_t j = 0; j < 64; j++) {
                const uint32_t row = q * 64 + j;
                if (row < a.n_rows) {
                    b0 |= (AV[(size_t)row * wpc] & 1) << j;
                    b1 |= (AM[(size_t)row * wpc] & 1) << j;
                }
            }
            ((uint64_t *)a.snap0)[q] = b0;
            ((uint64_t *)a.snap1)[q] = b1;
        }
    } else if (a.snap0 && blockIdx.x == 0) {  // lane 0 = bit 0 of the first lane-word, which CTA 0 holds
        for (uint32_t byte = tid; byte < a.snap_bytes; byte += nt) {
            uint32_t b0 = 0, b1 = 0;
            for (uint32_t j = 0; j < 8; j++) {
                const uint32_t row = byte * 8 + j;
                if (row < a.n_rows) {
                    b0 |= (uint32_t)(AV[(size_t)row * wpc] & 1) << j;
                    b1 |= (uint32_t)(AM[(size_t)row * wpc] & 1) << j;
                }
            }
            a.snap0[byte] = (uint8_t)b0;
            a.snap1[byte] = (uint8_t)b1;
        }
    }
    if (tid < wpc && w0 + tid < a.stride) {
        a.conflict[w0 + tid] = s_conflict[tid];
        a.unsettled[w0 + tid] = s_unsettled[tid];
        if (!a.host_flags) {
            if (s_conflict[tid]) atomicOr(&a.flags->any_conflict, 1u);
            if (s_unsettled[tid]) atomicOr(&a.flags->any_unsettled, 1u);
        }
    }
    if (a.host_flags) {  // one CTA: the flags of the run go straight to the host
        if (tid == 0) {
            uint64_t c = 0, u = 0;
            for (uint32_t j = 0; j < wpc; j++)
                if (w0 + j < a.stride) {
                    c |= s_conflict[j];
                    u |= s_unsettled[j];
                }
            RunFlags f;
            f.any_live = 0;
            f.any_conflict = c != 0;
            f.any_unsettled = u != 0;
            f.active = 0;
            a.host_flags[0] = f;
            *(unsigned long long *)(a.host_flags + 1) = apps;
        }
    } else if (tid == 0) {
        atomicMax(a.max_apps, apps);
    }
}

// ---------------------------------------------------------------------------- Engine: setup

Engine::Engine(Compiled &&net) : share_(std::make_shared<NetShare>(std::move(net))), net_(share_->net) {}
Engine::Engine(std::shared_ptr<NetShare> share) : share_(std::move(share)), net_(share_->net) {}

void Engine::copy_options_from(const Engine &o) {
    device_ = device_ready_ ? device_ : o.device_;
    use_resident_ = o.use_resident_;
    opt_elision_ = o.opt_elision_;
    opt_graph_ = o.opt_graph_;
    opt_frontier_ = o.opt_frontier_;
    opt_device_loop_ = o.opt_device_loop_;
    opt_lenient_ = o.opt_lenient_;
    opt_budget_inclusive_ = o.opt_budget_inclusive_;
    opt_u_ = o.opt_u_;
    opt_period2_ = o.opt_period2_;
    opt_l2_hints_ = o.opt_l2_hints_;
    opt_pdl_ = o.opt_pdl_;
    opt_discard_ = o.opt_discard_;
    opt_pair_ = o.opt_pair_;
    opt_hot_ = o.opt_hot_;
    opt_direct_ = o.opt_direct_;
    disc_short_ = o.disc_short_;
    disc_never_ = o.disc_never_;
    opt_flags_ = o.opt_flags_;
}

PairPlanDev::~PairPlanDev() {
    if (device < 0) return;
    cudaSetDevice(device);
    cudaFree(items);
    cudaFree(disc_rows);
}

DeviceNet::~DeviceNet() {
    if (device < 0) return;
    cudaSetDevice(device);
    cudaFree(fan0);
    cudaFree(fan1);
    cudaFree(fan2);
    cudaFree(kind2);
    cudaFree(far2);
    cudaFree(wide_off);
    cudaFree(wide_in);
    cudaFree(wide_kind);
    cudaFree(wide_nsel);
    cudaFree(wide_aux);
    cudaFree(res_off);
    cudaFree(res_in);
    cudaFree(levels);
}

template <typename Tv>
static int upload(Tv **dst, const std::vector<Tv> &src) {
    const size_t n = src.size() ? src.size() : 1;
    CU_TRY(cudaMalloc((void **)dst, n * sizeof(Tv)));
    if (src.size()) CU_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(Tv), cudaMemcpyHostToDevice));
    return B2_OK;
}

std::shared_ptr<DeviceNet> NetShare::acquire(int device, int &rc) {
    std::lock_guard<std::mutex> lock(m);
    rc = B2_OK;
    auto it = on_device.find(device);
    if (it != on_device.end())
        if (auto sp = it->second.lock()) return sp;
    auto d = std::make_shared<DeviceNet>();
    d->device = device;  // from here on the destructor frees whatever was allocated
    if ((rc = upload(&d->fan0, net.fan0)) || (rc = upload(&d->fan1, net.fan1)) || (rc = upload(&d->fan2, net.fan2)) ||
        (rc = upload(&d->kind2, net.kind2)) || (rc = upload(&d->far2, net.far2)) || (rc = upload(&d->wide_off, net.wide_off)) || (rc = upload(&d->wide_in, net.wide_in)) ||
        (rc = upload(&d->wide_kind, net.wide_kind)) || (rc = upload(&d->wide_nsel, net.wide_nsel)) || (rc = upload(&d->wide_aux, net.wide_aux)) ||
        (rc = upload(&d->res_off, net.res_off)) || (rc = upload(&d->res_in, net.res_in)) || (rc = upload(&d->levels, net.levels)))
        return nullptr;
    on_device[device] = d;
    return d;
}

// Frees whatever was allocated, whether or not ensure_device() ran to its end (a failure midway leaves
// device_ready_ false with some of the per-engine objects alive).
Engine::~Engine() {
    if (device_ < 0) return;
    if (cudaSetDevice(device_) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    if (stream_set_) cudaStreamSynchronize(ST);
    if (cin_stream_) cudaStreamSynchronize(CIN);
    if (cout_stream_) cudaStreamSynchronize(COUT);
    free_store();
    dnet_.reset();
    cudaFree(d_res_xslot_);
    cudaFree(d_flags_);
    cudaFree(d_pend_);
    cudaFree(d_loop_result_);
    if (h_loop_result_) cudaFreeHost(h_loop_result_);
    if (h_flags_) cudaFreeHost(h_flags_);
    if (h_snap_) cudaFreeHost(h_snap_);
    if (own_stream_ && stream_) cudaStreamDestroy(ST);
    if (cin_stream_) cudaStreamDestroy(CIN);
    if (cout_stream_) cudaStreamDestroy(COUT);
    if (side_stream_) {
        cudaStreamSynchronize((cudaStream_t)side_stream_);
        cudaStreamDestroy((cudaStream_t)side_stream_);
    }
    for (void *e : {ev_drive_consumed_, ev_drives_ready_, ev_gathered_, ev_out_free_, ev_settled_, ev_side_done_})
        if (e) cudaEventDestroy(EV(e));
    for (void *e : ev_flags_)
        if (e) cudaEventDestroy(EV(e));
    cudaGetLastError();
}

void Engine::free_store() {
    for (int i = 0; i < 2; i++) {
        cudaFree(val_[i]);
        cudaFree(vld_[i]);
        val_[i] = vld_[i] = nullptr;
    }
    cudaFree(drv_val_);
    cudaFree(drv_vld_);
    cudaFree(xdrv_val_);
    cudaFree(xdrv_vld_);
    cudaFree(d_extra_rows_);
    if (level_graph_exec_) {
        cudaGraphExecDestroy((cudaGraphExec_t)level_graph_exec_);
        level_graph_exec_ = nullptr;
    }
    cudaFree(d_rowchg_[0]);
    cudaFree(d_rowchg_[1]);
    d_rowchg_[0] = d_rowchg_[1] = nullptr;
    for (uint32_t *&p : d_act_list_) {
        cudaFree(p);
        p = nullptr;
    }
    frontier_valid_ = false;
    cudaFree(d_kv_);
    cudaFree(d_all_valid_);
    d_kv_ = nullptr;
    d_all_valid_ = nullptr;
    elided_ = false;
    cudaFree(d_changed_);
    cudaFree(d_conf_step_);
    cudaFree(d_live_);
    cudaFree(d_conflict_);
    cudaFree(d_unsettled_);
    for (int i = 0; i < 2; i++) {
        cudaFree(stage_val_[i]);
        cudaFree(stage_vld_[i]);
        stage_val_[i] = stage_vld_[i] = nullptr;
        stage_words_[i] = 0;
    }
    cin_pending_ = false;
    cudaFree(d_rows_);
    cudaFree(d_pack_);
    drv_val_ = drv_vld_ = xdrv_val_ = xdrv_vld_ = nullptr;
    d_extra_rows_ = nullptr;
    d_changed_ = d_conf_step_ = d_live_ = d_conflict_ = d_unsettled_ = nullptr;
    d_rows_ = nullptr;
    rows_cap_ = 0;
    rows_key_ = ptrs_key_ = 0;
    d_pack_ = nullptr;
    pack_cap_ = 0;
    extra_slot_.clear();
    extra_rows_.clear();
    extra_cap_ = 0;
    n_plain_extra_ = 0;
    snap_valid_ = false;
    pending_.clear();
    std::fill(res_xslot_.begin(), res_xslot_.end(), -1);
    res_xslot_dirty_ = true;
    store_bytes_ = 0;
}

int Engine::set_device(int device) {
    if (device_ready_) return device == device_ ? B2_OK : B2_ERR_INVALID_ARG;
    device_ = device;
    return B2_OK;
}

int Engine::set_stream(void *stream) {
    // The handle is used as given; 0 is CUDA's legacy default stream.  A private stream is
    // created only when no stream was ever set.
    if (device_ready_ && stream_set_ && stream != stream_) CU_TRY(cudaStreamSynchronize(ST));
    if (own_stream_ && stream_) {
        cudaStreamSynchronize(ST);
        cudaStreamDestroy(ST);
        own_stream_ = false;
    }
    stream_ = stream;
    stream_set_ = true;
    return B2_OK;
}

int Engine::ensure_device() {
    if (device_ready_) {
        CU_TRY(cudaSetDevice(device_));
        return B2_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_last_error(std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                       " (this engine has no CPU path)");
        cudaGetLastError();
        return B2_ERR_NO_DEVICE;
    }
    if (device_ < 0) {
        if (cudaGetDevice(&device_) != cudaSuccess) device_ = 0;
    }
    if (device_ >= count) {
        set_last_error("device ordinal out of range");
        return B2_ERR_INVALID_ARG;
    }
    CU_TRY(cudaSetDevice(device_));
    CU_TRY(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device_));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_list_[0], k_eval_list<2>, 256, 0));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_list_[1], k_eval_list<3>, 256, 0));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_list_[2], k_copy_list, 256, 0));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_list_[3], k_eval_wide_list, 256, 0));
    for (int &o : occ_list_) o = std::max(1, std::min(o, 4));
    {
        int coop = 0, occ = 0;
        CU_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device_));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_unit_delay_loop, 256, 0));
        loop_ctas_ = coop ? (uint32_t)(std::min(occ, 2) * sm_count_) : 0;
        if (!d_loop_result_) CU_TRY(cudaMalloc((void **)&d_loop_result_, sizeof(LoopResult) + 16 * sizeof(uint32_t)));
        if (!h_loop_result_) CU_TRY(cudaMallocHost((void **)&h_loop_result_, sizeof(LoopResult)));
    }
    // the kernels are built for sm_100a only: fail loudly on anything else
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k_clear_flags);
    if (e != cudaSuccess) {
        set_last_error(std::string("sm_100a kernels not loadable on this device: ") + cudaGetErrorString(e));
        cudaGetLastError();
        return B2_ERR_NO_DEVICE;
    }
    if (!stream_set_) {
        cudaStream_t s;
        CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        stream_ = s;
        own_stream_ = true;
        stream_set_ = true;
    }
    int rc;
    if (!dnet_) {  // one copy of the netlist per device, shared with every replica of this build
        dnet_ = share_->acquire(device_, rc);
        if (!dnet_) return rc;
        d_fan0_ = dnet_->fan0; d_fan1_ = dnet_->fan1; d_fan2_ = dnet_->fan2; d_kind2_ = dnet_->kind2; d_far2_ = dnet_->far2;
        d_wide_off_ = dnet_->wide_off; d_wide_in_ = dnet_->wide_in; d_wide_kind_ = dnet_->wide_kind; d_wide_nsel_ = dnet_->wide_nsel; d_wide_aux_ = dnet_->wide_aux;
        d_res_off_ = dnet_->res_off; d_res_in_ = dnet_->res_in; d_levels_ = dnet_->levels;
    }
    // a retry after a failure midway finds the earlier objects and keeps them
    if (!d_res_xslot_) {
        res_xslot_.assign(net_.n_res(), -1);
        if ((rc = upload(&d_res_xslot_, res_xslot_))) return rc;
        res_xslot_dirty_ = false;
    }
    if (!d_flags_) CU_TRY(cudaMalloc((void **)&d_flags_, 8 * sizeof(RunFlags)));
    CU_TRY(cudaMemset(d_flags_, 0, 8 * sizeof(RunFlags)));
    if (!h_flags_) {
        CU_TRY(cudaHostAlloc((void **)&h_flags_, 8 * sizeof(RunFlags), cudaHostAllocMapped));
        if (cudaHostGetDevicePointer(&h_flags_dev_, h_flags_, 0) != cudaSuccess) {
            cudaGetLastError();
            h_flags_dev_ = nullptr;  // no direct stores into host memory: flags and snapshot are copied
        }
    }
    for (int i = 0; i < 8; i++) {
        if (ev_flags_[i]) continue;
        cudaEvent_t ev;
        CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ev_flags_[i] = ev;
    }
    for (void **st : {&cin_stream_, &cout_stream_}) {
        if (*st) continue;
        cudaStream_t a;
        CU_TRY(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        *st = a;
    }
    for (void **e : {&ev_drive_consumed_, &ev_drives_ready_, &ev_gathered_, &ev_out_free_}) {
        if (*e) continue;
        cudaEvent_t ev;
        CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        *e = ev;
    }
    device_ready_ = true;
    return B2_OK;
}

int Engine::synchronize() {
    if (!device_ready_) return B2_OK;
    CU_TRY(cudaSetDevice(device_));
    CU_TRY(cudaStreamSynchronize(CIN));
    CU_TRY(cudaStreamSynchronize(ST));
    CU_TRY(cudaStreamSynchronize(COUT));
    return B2_OK;
}

// The drive store is written by the copy-in stream (b2_set_drives_lanes) and read/written by the
// main stream; two events order the two streams without serialising whole tiles.
int Engine::drives_begin() {
    if (cin_pending_) {
        CU_TRY(cudaStreamWaitEvent(ST, EV(ev_drives_ready_), 0));
        cin_pending_ = false;
    }
    return B2_OK;
}
int Engine::drives_end() {
    CU_TRY(cudaEventRecord(EV(ev_drive_consumed_), ST));
    return B2_OK;
}

static uint32_t stride_for(uint32_t T) {
    if (T <= 2) return 2;
    if (T < 32) return (T + 1) & ~1u;
    return (T + 15) & ~15u;  // 128-byte aligned rows
}

static uint64_t bytes_per_lane_word(const Compiled &n, bool second) {
    // A (+B) planes, drive rows, status words
    return ((uint64_t)n.n_rows * (second ? 2 : 1) + n.n_src) * 16 + 5 * 8;
}

uint64_t Engine::max_lanes(uint64_t bytes) const {
    if (bytes == 0) {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        bytes = fr;
    }
    const uint64_t per = bytes_per_lane_word(net_, net_.cyclic);
    uint64_t words = bytes / (per ? per : 1);
    if (words > 32) words &= ~15ull;
    if (words > 0x7FFFFFF0ull) words = 0x7FFFFFF0ull;
    return words * 64;
}

int Engine::set_lanes(uint64_t n_lanes) {
    if (n_lanes == 0 || n_lanes > 64ull * 0x7FFFFFF0ull) return B2_ERR_INVALID_ARG;
    int rc = ensure_device();
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(CIN));
    CU_TRY(cudaStreamSynchronize(ST));
    CU_TRY(cudaStreamSynchronize(COUT));
    free_store();
    n_lanes_ = n_lanes;
    T_ = (uint32_t)((n_lanes + 63) / 64);
    stride_ = stride_for(T_);
    const size_t row_words = (size_t)(net_.n_rows ? net_.n_rows : 1) * stride_;
    const size_t src_words = (size_t)(net_.n_src ? net_.n_src : 1) * stride_;
    auto alloc0 = [&](uint64_t **p, size_t words) -> int {
        CU_TRY(cudaMalloc((void **)p, words * 8));
        CU_TRY(cudaMemsetAsync(*p, 0, words * 8, ST));
        store_bytes_ += words * 8;
        return B2_OK;
    };
    if ((rc = alloc0(&val_[0], row_words)) || (rc = alloc0(&vld_[0], row_words)) || (rc = alloc0(&drv_val_, src_words)) ||
        (rc = alloc0(&drv_vld_, src_words)) || (rc = alloc0(&d_changed_, stride_)) || (rc = alloc0(&d_conf_step_, stride_)) ||
        (rc = alloc0(&d_live_, stride_)) || (rc = alloc0(&d_conflict_, stride_)) || (rc = alloc0(&d_unsettled_, stride_))) {
        free_store();
        n_lanes_ = 0;
        T_ = stride_ = 0;
        return rc;
    }
    auto alloc_flags = [&]() -> int {
        CU_TRY(cudaMalloc((void **)&d_kv_, net_.n_rows ? net_.n_rows : 1));
        CU_TRY(cudaMemsetAsync(d_kv_, 0, net_.n_rows ? net_.n_rows : 1, ST));
        CU_TRY(cudaMalloc((void **)&d_all_valid_, 4));
        CU_TRY(cudaMemsetAsync(d_all_valid_, 0, 4, ST));
        return B2_OK;
    };
    if ((rc = alloc_flags())) {  // unwind: no half-built store stays behind
        free_store();
        n_lanes_ = 0;
        T_ = stride_ = 0;
        return rc;
    }
    store_bytes_ += net_.n_rows;
    // A.6-5 alternative: wires read Undefined (1, 0) instead of High-Z (0, 0) until the first run
    if (net_.sem_flags & SEM_INIT_X) CU_TRY(cudaMemsetAsync(val_[0], 0xFF, row_words * 8, ST));
    cur_ = 0;
    cold_ = true;
    last_mode_ = 0;
    last_apps_ = 0;
    CU_TRY(cudaStreamSynchronize(ST));  // the zero-fills are done before any other stream touches the store
    return B2_OK;
}

int Engine::ensure_lanes() {
    int rc = ensure_device();
    if (rc) return rc;
    if (n_lanes_ == 0) return set_lanes(1);
    return B2_OK;
}

int Engine::ensure_second_buffer() {
    if (val_[1]) return B2_OK;
    const size_t row_words = (size_t)(net_.n_rows ? net_.n_rows : 1) * stride_;
    CU_TRY(cudaMalloc((void **)&val_[1], row_words * 8));
    CU_TRY(cudaMalloc((void **)&vld_[1], row_words * 8));
    CU_TRY(cudaMemsetAsync(val_[1], 0, row_words * 8, ST));
    CU_TRY(cudaMemsetAsync(vld_[1], 0, row_words * 8, ST));
    store_bytes_ += row_words * 16;
    return B2_OK;
}

int Engine::ensure_stage(bool out, size_t words) {
    const int i = out ? 1 : 0;
    if (stage_words_[i] >= words) return B2_OK;
    CU_TRY(cudaStreamSynchronize(ST));
    CU_TRY(cudaStreamSynchronize(out ? COUT : CIN));
    cudaFree(stage_val_[i]);
    cudaFree(stage_vld_[i]);
    stage_val_[i] = stage_vld_[i] = nullptr;
    store_bytes_ -= stage_words_[i] * 16;
    stage_words_[i] = 0;
    CU_TRY(cudaMalloc((void **)&stage_val_[i], words * 8));
    CU_TRY(cudaMalloc((void **)&stage_vld_[i], words * 8));
    stage_words_[i] = words;
    store_bytes_ += words * 16;
    return B2_OK;
}

int Engine::ensure_rows_buffer(uint32_t n) {
    // holds n row indices and, behind them (8-byte aligned), two tables of n row pointers
    if (rows_cap_ >= n) return B2_OK;
    CU_TRY(cudaStreamSynchronize(ST));
    CU_TRY(cudaStreamSynchronize(CIN));  // a scatter kernel may still be reading the pointer tables
    cudaFree(d_rows_);
    d_rows_ = nullptr;
    rows_cap_ = 0;
    const uint32_t cap = std::max<uint32_t>(n, 1024);
    const size_t idx_bytes = (((size_t)cap * 4) + 15) & ~(size_t)15;
    CU_TRY(cudaMalloc((void **)&d_rows_, idx_bytes + (size_t)cap * 16));
    rows_cap_ = cap;
    return B2_OK;
}

// Row pointers of the drive of a bit-net; gate-driven bits get a slot in the extra store.
int Engine::drive_target(uint32_t bitnet, uint64_t **val, uint64_t **vld) {
    if (!net_.bit_driven[bitnet]) {
        const size_t o = (size_t)net_.row_of_bit[bitnet] * stride_;
        *val = drv_val_ + o;
        *vld = drv_vld_ + o;
        return B2_OK;
    }
    auto it = extra_slot_.find(bitnet);
    uint32_t slot;
    if (it != extra_slot_.end()) {
        slot = it->second;
    } else {
        slot = (uint32_t)extra_rows_.size();
        if (slot >= extra_cap_) {
            const uint32_t cap = extra_cap_ ? extra_cap_ * 2 : 64;
            uint64_t *nv = nullptr, *nm = nullptr;
            uint32_t *nr = nullptr;
            auto grow = [&]() -> int {
                CU_TRY(cudaMalloc((void **)&nv, (size_t)cap * stride_ * 8));
                CU_TRY(cudaMalloc((void **)&nm, (size_t)cap * stride_ * 8));
                CU_TRY(cudaMalloc((void **)&nr, (size_t)cap * 4));
                CU_TRY(cudaMemsetAsync(nv, 0, (size_t)cap * stride_ * 8, ST));
                CU_TRY(cudaMemsetAsync(nm, 0, (size_t)cap * stride_ * 8, ST));
                if (extra_cap_) {
                    CU_TRY(cudaMemcpyAsync(nv, xdrv_val_, (size_t)extra_cap_ * stride_ * 8, cudaMemcpyDeviceToDevice, ST));
                    CU_TRY(cudaMemcpyAsync(nm, xdrv_vld_, (size_t)extra_cap_ * stride_ * 8, cudaMemcpyDeviceToDevice, ST));
                }
                CU_TRY(cudaStreamSynchronize(ST));
                return B2_OK;
            };
            if (int grc = grow()) {  // the old store stays as it was
                cudaFree(nv);
                cudaFree(nm);
                cudaFree(nr);
                return grc;
            }
            cudaFree(xdrv_val_);
            cudaFree(xdrv_vld_);
            cudaFree(d_extra_rows_);
            d_extra_rows_ = nr;
            extra_rows_dirty_ = true;
            xdrv_val_ = nv;
            xdrv_vld_ = nm;
            ptrs_key_ = 0;
            store_bytes_ += (size_t)(cap - extra_cap_) * stride_ * 16;
            extra_cap_ = cap;
        }
        extra_slot_[bitnet] = slot;
        const int32_t res = net_.res_of_bit[bitnet];
        if (res >= 0) {  // tri-state net: the drive is an operand of its resolve, not a merge into a gate row
            extra_rows_.push_back(0xFFFFFFFFu);
            res_xslot_[res] = (int32_t)slot;
            res_xslot_dirty_ = true;
        } else {
            extra_rows_.push_back(net_.row_of_bit[bitnet]);
            n_plain_extra_++;
        }
        extra_rows_dirty_ = true;
    }
    *val = xdrv_val_ + (size_t)slot * stride_;
    *vld = xdrv_vld_ + (size_t)slot * stride_;
    return B2_OK;
}

int Engine::sync_extra_tables() {
    if (extra_rows_dirty_ && !extra_rows_.empty()) {
        CU_TRY(cudaMemcpyAsync(d_extra_rows_, extra_rows_.data(), extra_rows_.size() * 4, cudaMemcpyHostToDevice, ST));
        extra_rows_dirty_ = false;
    }
    if (res_xslot_dirty_ && !res_xslot_.empty()) {
        CU_TRY(cudaMemcpyAsync(d_res_xslot_, res_xslot_.data(), res_xslot_.size() * 4, cudaMemcpyHostToDevice, ST));
        res_xslot_dirty_ = false;
    }
    return B2_OK;
}

// ---------------------------------------------------------------------------- Engine: drives

// gsim's set_wire_drive is a plain memory write and the reference calls it once per input per eval
// (crates/digilogic_netcode/src/server.rs:447-451), so the call itself does no CUDA work: drives are
// queued on the host and reach the drive store in one upload + one launch before the next run (or before
// any other call that writes drives, to keep the call order).
int Engine::set_wire_drive(uint32_t wire, const uint32_t *p0, const uint32_t *p1, size_t words) {
    if (wire >= net_.n_wires()) return B2_ERR_INVALID_WIRE_ID;
    if (words && (!p0 || !p1)) return B2_ERR_NULL;
    if (!device_ready_ || n_lanes_ == 0) {  // first use: device and store; afterwards the call is pure host work
        int rc = ensure_lanes();
        if (rc) return rc;
    }
    PendingDrive d;
    d.wire = wire;
    for (uint32_t i = 0; i < 8; i++) {
        d.p0[i] = i < words ? p0[i] : 0;
        d.p1[i] = i < words ? p1[i] : 0;
    }
    pending_.push_back(d);
    if (pending_.size() >= (1u << 16)) return flush_pending_drives();
    return B2_OK;
}

// The queued drives as at most `cap` (row pointers, planes) entries for the resident kernel's arguments; *n = 0 (and the queue
// untouched) when they do not fit.
int Engine::pending_to_bits(void *out, uint32_t cap, uint32_t *n) {
    *n = 0;
    size_t n_bits = 0;
    for (const PendingDrive &d : pending_) n_bits += net_.width[d.wire];
    if (n_bits == 0 || n_bits > cap) return B2_OK;
    // a wire driven twice keeps only its last drive: later entries overwrite earlier ones in queue order (one thread block
    // writes them, but in an unspecified order) — so drop the earlier ones here
    std::unordered_map<uint32_t, size_t> last;
    for (size_t i = 0; i < pending_.size(); i++) last[pending_[i].wire] = i;
    PendingBit *bits = (PendingBit *)out;
    int rc;
    for (int pass = 0; pass < 2; pass++) {  // two passes: the first may grow the extra store and move earlier rows
        size_t k = 0;
        for (size_t i = 0; i < pending_.size(); i++) {
            const PendingDrive &d = pending_[i];
            if (last[d.wire] != i) continue;
            for (uint32_t j = 0; j < net_.width[d.wire]; j++, k++) {
                PendingBit &b = bits[k];
                if ((rc = drive_target(net_.bit_off[d.wire] + j, &b.val, &b.vld))) return rc;
                b.flags = ((d.p0[j >> 5] >> (j & 31)) & 1) | (((d.p1[j >> 5] >> (j & 31)) & 1) << 1);
                b.pad = 0;
            }
        }
        *n = (uint32_t)k;
    }
    pending_.clear();
    return B2_OK;
}

int Engine::flush_pending_drives() {
    if (pending_.empty()) return B2_OK;
    int rc;
    // one launch writes all queued bits concurrently, so a wire driven twice keeps only its last drive
    if (pending_.size() > 1) {
        std::unordered_map<uint32_t, size_t> last;
        for (size_t i = 0; i < pending_.size(); i++) last[pending_[i].wire] = i;
        if (last.size() != pending_.size()) {
            size_t k = 0;
            for (size_t i = 0; i < pending_.size(); i++)
                if (last[pending_[i].wire] == i) pending_[k++] = pending_[i];
            pending_.resize(k);
        }
    }
    size_t n_bits = 0;
    for (const PendingDrive &d : pending_) n_bits += net_.width[d.wire];
    // two passes: the first may grow the extra store (drives on gate-driven nets) and move earlier rows
    std::vector<PendingBit> bits(n_bits);
    for (int pass = 0; pass < 2; pass++) {
        size_t k = 0;
        for (const PendingDrive &d : pending_)
            for (uint32_t j = 0; j < net_.width[d.wire]; j++, k++) {
                PendingBit &b = bits[k];
                if ((rc = drive_target(net_.bit_off[d.wire] + j, &b.val, &b.vld))) return rc;
                b.flags = ((d.p0[j >> 5] >> (j & 31)) & 1) | (((d.p1[j >> 5] >> (j & 31)) & 1) << 1);
                b.pad = 0;
            }
    }
    pending_.clear();
    if (pend_cap_ < n_bits) {
        CU_TRY(cudaStreamSynchronize(ST));
        cudaFree(d_pend_);
        d_pend_ = nullptr;
        pend_cap_ = 0;
        const size_t cap = std::max<size_t>(n_bits, 256);
        CU_TRY(cudaMalloc((void **)&d_pend_, cap * sizeof(PendingBit)));
        pend_cap_ = cap;
    }
    // pageable source: the call returns once the data is staged, so `bits` may go out of scope
    CU_TRY(cudaMemcpyAsync(d_pend_, bits.data(), n_bits * sizeof(PendingBit), cudaMemcpyHostToDevice, ST));
    if ((rc = drives_begin())) return rc;
    const uint32_t word_blocks = (stride_ + 255) / 256;
    for (size_t done = 0; done < n_bits;) {  // grid.x is bounded; one launch covers any realistic flush
        const size_t chunk = std::min<size_t>(n_bits - done, 0x7FFFFFFFull / word_blocks);
        k_broadcast_bits<<<(uint32_t)(chunk * word_blocks), 256, 0, ST>>>((const PendingBit *)d_pend_ + done, (uint32_t)chunk, stride_,
                                                                       word_blocks);
        LAUNCHED();
        done += chunk;
    }
    return drives_end();
}

// Lane 0 of every row, bit-packed, on the host: what get_wire_state serves from.  The reference reads
// every net back after each eval, one call per net (crates/digilogic_netcode/src/server.rs:375-383); with
// the snapshot that is one kernel and one copy per eval instead of a device round trip per net.
int Engine::refresh_snapshot() {
    if (snap_valid_) return B2_OK;
    int rc = materialise_valid();
    if (rc) return rc;
    const size_t n_bytes = (size_t)net_.n_rows / 8 + 1;
    if (pack_cap_ < 2 * n_bytes) {
        CU_TRY(cudaStreamSynchronize(ST));
        cudaFree(d_pack_);
        d_pack_ = nullptr;
        pack_cap_ = 0;
        CU_TRY(cudaMalloc((void **)&d_pack_, 2 * n_bytes + 64));
        pack_cap_ = 2 * n_bytes + 64;
    }
    snap_.resize(2 * n_bytes);
    k_pack_lane0<<<(uint32_t)((n_bytes + 255) / 256), 256, 0, ST>>>(nullptr, net_.n_rows, val_[cur_], vld_[cur_], stride_, d_pack_,
                                                                    d_pack_ + n_bytes, (uint32_t)n_bytes);
    LAUNCHED();
    CU_TRY(cudaMemcpyAsync(snap_.data(), d_pack_, 2 * n_bytes, cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaStreamSynchronize(ST));
    snap_valid_ = true;
    return B2_OK;
}

int Engine::set_wire_drive_lanes(uint32_t wire, uint32_t bit, uint32_t w0, uint32_t nw, const uint64_t *value,
                                 const uint64_t *valid) {
    if (wire >= net_.n_wires() || bit >= net_.width[wire]) return B2_ERR_INVALID_WIRE_ID;
    if (!value || !valid) return B2_ERR_NULL;
    int rc = ensure_lanes();
    if (rc) return rc;
    if ((uint64_t)w0 + nw > T_) return B2_ERR_RANGE;
    if ((rc = flush_pending_drives())) return rc;
    uint64_t *dv, *dm;
    if ((rc = drive_target(net_.bit_off[wire] + bit, &dv, &dm))) return rc;
    if ((rc = drives_begin())) return rc;
    CU_TRY(cudaMemcpyAsync(dv + w0, value, (size_t)nw * 8, cudaMemcpyHostToDevice, ST));
    CU_TRY(cudaMemcpyAsync(dm + w0, valid, (size_t)nw * 8, cudaMemcpyHostToDevice, ST));
    if ((rc = drives_end())) return rc;
    CU_TRY(cudaStreamSynchronize(ST));  // caller may reuse its buffers
    return B2_OK;
}

static uint64_t hash_wires(const uint32_t *wires, uint32_t n) {
    uint64_t h = 0xCBF29CE484222325ull ^ n;
    for (uint32_t i = 0; i < n; i++) h = (h ^ wires[i]) * 0x100000001B3ull;
    return h | 1;  // never 0 (0 = nothing cached)
}

// Uploads row indices (store rows) or drive row pointer tables for bit 0 of `wires`.  The last
// list of each kind stays cached on the device, so a tile loop re-using the same input and
// output lists neither re-uploads nor synchronises.
int Engine::upload_rows(const uint32_t *wires, uint32_t n, bool drive_rows) {
    const uint32_t cap_before = rows_cap_;
    int rc = ensure_rows_buffer(n);
    if (rc) return rc;
    if (rows_cap_ != cap_before) rows_key_ = ptrs_key_ = 0;
    for (uint32_t i = 0; i < n; i++)
        if (wires[i] >= net_.n_wires()) return B2_ERR_INVALID_WIRE_ID;
    const uint64_t key = hash_wires(wires, n);
    // a hash hit counts only if the cached list is the same list (two lists can share a hash)
    std::vector<uint32_t> &cached = drive_rows ? ptrs_wires_ : rows_wires_;
    const bool same = cached.size() == n && (n == 0 || std::memcmp(cached.data(), wires, (size_t)n * 4) == 0);
    if (!same) {
        if (drive_rows) ptrs_key_ = 0;
        else rows_key_ = 0;
        cached.assign(wires, wires + n);
    }
    if (!drive_rows) {
        if (rows_key_ == key) return B2_OK;
        std::vector<uint32_t> rows(n);
        for (uint32_t i = 0; i < n; i++) rows[i] = net_.row_of_bit[net_.bit_off[wires[i]]];
        CU_TRY(cudaMemcpyAsync(d_rows_, rows.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ST));
        CU_TRY(cudaStreamSynchronize(ST));
        rows_key_ = key;
    } else {
        if (ptrs_key_ == key) return B2_OK;
        CU_TRY(cudaStreamSynchronize(CIN));  // a scatter kernel may still be reading the old tables
        std::vector<uint64_t *> tv(n), tm(n);
        // two passes: the first may grow the extra store and move earlier rows
        for (int pass = 0; pass < 2; pass++)
            for (uint32_t i = 0; i < n; i++)
                if ((rc = drive_target(net_.bit_off[wires[i]], &tv[i], &tm[i]))) return rc;
        const size_t idx_bytes = (((size_t)rows_cap_ * 4) + 15) & ~(size_t)15;
        uint64_t **d_tv = (uint64_t **)((char *)d_rows_ + idx_bytes), **d_tm = d_tv + rows_cap_;
        CU_TRY(cudaMemcpyAsync(d_tv, tv.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ST));
        CU_TRY(cudaMemcpyAsync(d_tm, tm.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ST));
        CU_TRY(cudaStreamSynchronize(ST));
        ptrs_key_ = key;
    }
    return B2_OK;
}

int Engine::set_drives_lanes(const uint32_t *wires, uint32_t n, const uint64_t *value, const uint64_t *valid, size_t src_stride) {
    if (n && (!wires || !value || !valid)) return B2_ERR_NULL;
    int rc = ensure_lanes();
    if (rc) return rc;
    if (n == 0) return B2_OK;
    if (src_stride < T_) return B2_ERR_RANGE;
    if ((rc = flush_pending_drives())) return rc;
    if ((rc = upload_rows(wires, n, true))) return rc;
    if ((rc = ensure_stage(false, (size_t)n * T_))) return rc;
    // copy-in stream: wait until the main stream has consumed the previous drives (early in the
    // previous settle), then H2D + scatter while that settle is still running
    CU_TRY(cudaStreamWaitEvent(CIN, EV(ev_drive_consumed_), 0));
    // (the planes may also be device memory: the batch runner passes either)
    CU_TRY(cudaMemcpy2DAsync(stage_val_[0], (size_t)T_ * 8, value, src_stride * 8, (size_t)T_ * 8, n, cudaMemcpyDefault, CIN));
    CU_TRY(cudaMemcpy2DAsync(stage_vld_[0], (size_t)T_ * 8, valid, src_stride * 8, (size_t)T_ * 8, n, cudaMemcpyDefault, CIN));
    const size_t idx_bytes = (((size_t)rows_cap_ * 4) + 15) & ~(size_t)15;
    uint64_t **d_tv = (uint64_t **)((char *)d_rows_ + idx_bytes), **d_tm = d_tv + rows_cap_;
    if (T_ % 2 == 0) {  // rows, staging planes and T_ are all even: 128-bit accesses
        dim3 block(64, 4), grid((T_ / 2 + 63) / 64, (n + 3) / 4);
        k_scatter_rows<true><<<grid, block, 0, CIN>>>(d_tv, d_tm, n, stage_val_[0], stage_vld_[0], T_, T_);
    } else {
        dim3 block(64, 4), grid((T_ + 63) / 64, (n + 3) / 4);
        k_scatter_rows<false><<<grid, block, 0, CIN>>>(d_tv, d_tm, n, stage_val_[0], stage_vld_[0], T_, T_);
    }
    LAUNCHED();
    CU_TRY(cudaEventRecord(EV(ev_drives_ready_), CIN));
    cin_pending_ = true;
    CU_TRY(cudaStreamSynchronize(CIN));  // the caller may reuse its buffers; the main stream is not waited for
    return B2_OK;
}

int Engine::fill_stimulus(const uint32_t *wires, uint32_t n, uint64_t seed, uint64_t word_base, int mode) {
    if (n && !wires) return B2_ERR_NULL;
    if (mode != 0 && mode != 1) return B2_ERR_INVALID_ARG;
    int rc = ensure_lanes();
    if (rc) return rc;
    if (n == 0) return B2_OK;
    if ((rc = flush_pending_drives())) return rc;
    if ((rc = upload_rows(wires, n, true))) return rc;
    const size_t idx_bytes = (((size_t)rows_cap_ * 4) + 15) & ~(size_t)15;
    uint64_t **d_tv = (uint64_t **)((char *)d_rows_ + idx_bytes), **d_tm = d_tv + rows_cap_;
    if ((rc = drives_begin())) return rc;
    dim3 block(64, 4), grid((stride_ / 2 + 63) / 64, (n + 3) / 4);  // stride_ is even
    k_fill_stimulus<<<grid, block, 0, ST>>>(d_tv, d_tm, n, stride_, seed, word_base, mode);
    LAUNCHED();
    if ((rc = drives_end())) return rc;
    return B2_OK;
}

// ---------------------------------------------------------------------------- Engine: settle loops

namespace {
struct Shape {
    dim3 block;
    uint32_t pair_blocks;
};
// BX = power of two covering the row (<= 256), BY = 256 / BX gates per CTA slice.
Shape shape_for(uint32_t pairs) {
    uint32_t bx = 1;
    while (bx < pairs && bx < 256) bx <<= 1;
    Shape s;
    s.block = dim3(bx, 256 / bx);
    s.pair_blocks = (pairs + bx - 1) / bx;
    return s;
}
}  // namespace

// Elided valid planes are expanded before any path that reads the store as plain two-plane rows.
int Engine::materialise_valid() {
    if (!elided_) return B2_OK;
    const uint32_t pairs = stride_ / 2;
    uint32_t bx = 1;
    while (bx < pairs && bx < 256) bx <<= 1;
    const dim3 block(bx, 256 / bx);
    const uint32_t pair_blocks = (pairs + bx - 1) / bx, rb = (net_.n_rows + block.y - 1) / block.y;
    k_materialise_valid<<<pair_blocks * rb, block, 0, ST>>>(vld_[cur_], d_kv_, d_all_valid_, net_.n_rows, stride_, pairs, pair_blocks);
    LAUNCHED();
    elided_ = false;
    frontier_valid_ = false;
    return B2_OK;
}

#define EVAL2_U 4

// The level-pair plan for this netlist, device and set of kept rows: taken from the netlist's cache when another replica
// has built it, else planned on the host (plan_pairs) and uploaded.  pair_ stays null when the netlist is not eligible.
int Engine::ensure_pair_plan() {
    if (pair_ || pair_tried_) return B2_OK;
    pair_tried_ = true;
    std::lock_guard<std::mutex> lock(share_->m);
    auto it = share_->pair_plans.find(device_);
    if (it != share_->pair_plans.end()) {
        std::shared_ptr<PairPlanDev> have = it->second.lock();
        if (have && have->keep_all == !opt_discard_ && (have->keep_all || have->keep_rows == keep_rows_)) {
            pair_ = have;
            return B2_OK;
        }
    }
    PairPlanHost plan;
    if (!plan_pairs(net_, keep_rows_, plan, !opt_discard_)) return B2_OK;
    auto dev = std::make_shared<PairPlanDev>();
    dev->device = device_;
    dev->keep_rows = keep_rows_;
    dev->keep_all = !opt_discard_;
    dev->n_items = plan.items.size();
    dev->n_children = plan.n_children;
    dev->n_unstored = plan.n_unstored;
    for (const PairLaunch &l : plan.launches) {
        dev->launch_tab.push_back(l.item_begin);
        dev->launch_tab.push_back(l.item_end);
        dev->launch_tab.push_back(l.disc_begin);
        dev->launch_tab.push_back(l.disc_end);
    }
    CU_TRY(cudaMalloc(&dev->items, std::max<size_t>(plan.items.size(), 1) * sizeof(PairItem)));
    CU_TRY(cudaMalloc((void **)&dev->disc_rows, std::max<size_t>(plan.disc_rows.size(), 1) * 4));
    CU_TRY(cudaMemcpyAsync(dev->items, plan.items.data(), plan.items.size() * sizeof(PairItem), cudaMemcpyHostToDevice, ST));
    CU_TRY(cudaMemcpyAsync(dev->disc_rows, plan.disc_rows.data(), plan.disc_rows.size() * 4, cudaMemcpyHostToDevice, ST));
    CU_TRY(cudaStreamSynchronize(ST));  // the host vectors go away; other replicas' streams may use the plan next
    share_->pair_plans[device_] = dev;
    pair_ = dev;
    return B2_OK;
}

// The (at most four) rows with the largest fan-out among the fast gates, if a hundred gates or more read them: staged in
// shared memory by k_eval_list.
void Engine::find_hot_rows() {
    if (hot_ready_) return;
    hot_ready_ = true;
    n_hot_ = 0;
    std::vector<uint32_t> fanout(net_.n_rows, 0);
    for (uint32_t g = 0; g < net_.n_fast(); g++) {
        const uint32_t f[3] = {net_.fan0[g], net_.fan1[g], net_.fan2[g]};
        fanout[f[0]]++;
        if (f[1] != f[0]) fanout[f[1]]++;
        if (f[2] != f[0] && f[2] != f[1]) fanout[f[2]]++;
    }
    for (uint32_t h = 0; h < kHotRows; h++) {
        uint32_t best = 0, best_n = kHotMinFanout - 1;
        for (uint32_t r = 0; r < net_.n_rows; r++)
            if (fanout[r] > best_n) {
                best = r;
                best_n = fanout[r];
            }
        if (best_n < kHotMinFanout) break;
        hot_rows_[n_hot_++] = best;
        fanout[best] = 0;
    }
}

void Engine::hot_rows(uint32_t *rows, uint32_t *n) {
    find_hot_rows();
    *n = opt_hot_ ? n_hot_ : 0;
    for (uint32_t h = 0; h < *n; h++) rows[h] = hot_rows_[h];
}

void Engine::pair_info(uint64_t *launches, uint64_t *children, uint64_t *unstored) const {
    const bool on = pair_ && last_mode_ == 1 && last_pair_;
    if (launches) *launches = on ? pair_->launch_tab.size() / 4 : 0;
    if (children) *children = on ? pair_->n_children : 0;
    if (unstored) *unstored = on ? pair_->n_unstored : 0;
}
static const uint32_t kFlagSlots = 8, kFlagLag = 2;
// device-side loop: entered when listed gates x row words of an application is at most this, left above four times this
static const uint64_t kLoopEntryRowWords = 1ull << 20;

int Engine::run_levelised(bool async) {
    const uint32_t pairs = stride_ / 2;
    const Shape sh = shape_for(pairs);
    uint64_t *V = val_[cur_], *M = vld_[cur_];
    int rc0 = drives_begin();
    if (rc0) return rc0;
    // valid-plane elision: only when no gate output carries an external drive (those merge into valid planes)
    const bool kv = opt_elision_ && extra_rows_.empty();
    if (kv && net_.n_src) CU_TRY(cudaMemsetAsync(d_kv_, 1, net_.n_src, ST));
    if (net_.n_src) {
        const uint32_t rb = (net_.n_src + sh.block.y - 1) / sh.block.y;
        k_apply_drive<false><<<sh.pair_blocks * rb, sh.block, 0, ST>>>(drv_val_, drv_vld_, nullptr, nullptr, V, M, nullptr,
                                                                        net_.n_src, stride_, pairs, sh.pair_blocks,
                                                                        kv ? d_kv_ : nullptr, T_);
        LAUNCHED();
    }
    if (kv) {
        k_all_valid<<<1, 256, 0, ST>>>(d_kv_, net_.n_src, d_all_valid_);
        LAUNCHED();
    }
    if (extra_rows_.empty() && (rc0 = drives_end())) return rc0;  // the drive rows may be refilled from here on
    // The level sweep is a fixed sequence of launches for a given store: it is captured once into a
    // CUDA graph and replayed per run, which takes the per-launch CPU cost and most of the
    // launch-to-launch gap out of small tiles (hundreds of levels of a few microseconds each).
    const bool want_graph = opt_graph_ && net_.levels.size() >= 8;
    const uint64_t gkey = ((uint64_t)(uintptr_t)V) ^ ((uint64_t)stride_ << 2) ^ (kv ? 1u : 0u) ^ (opt_l2_hints_ ? 2u : 0u) ^ (opt_discard_ ? 8u : 0u) ^ ((uint64_t)opt_pdl_ << 4) ^ ((uint64_t)opt_pair_ << 6);
    // level-pair fusion: only where rows need not survive their last reader (set_discard), plain two-plane rows, whole-line rows
    const bool pair_wanted = opt_pair_ == 1 || (opt_pair_ == 2 && net_.depth && net_.n_g2() / net_.depth <= kPairAutoWidth);
    // (with set_discard: rows die and are dropped, which needs whole-line rows; without: every row is kept, only the launches are fused)
    const bool pair_path = pair_wanted && !kv && extra_rows_.empty() && (!opt_discard_ || stride_ % 16 == 0);
    if (pair_path && (rc0 = ensure_pair_plan())) return rc0;
    const bool use_pair = pair_path && pair_ != nullptr;
    last_pair_ = use_pair;
    if (want_graph && level_graph_exec_ && level_graph_key_ == gkey) {
        CU_TRY(cudaGraphLaunch((cudaGraphExec_t)level_graph_exec_, ST));
        g_launches.fetch_add(level_graph_nodes_, std::memory_order_relaxed);
    } else {
        bool capturing = false;
        if (want_graph) {
            if (level_graph_exec_) {
                cudaGraphExecDestroy((cudaGraphExec_t)level_graph_exec_);
                level_graph_exec_ = nullptr;
            }
            capturing = cudaStreamBeginCapture(ST, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (!capturing) cudaGetLastError();
        }
        uint32_t nodes = 0;
        if (use_pair) {
            const std::vector<uint32_t> &tab = pair_->launch_tab;
            for (size_t l = 0; l + 3 < tab.size(); l += 4) {
                PairArgs a;
                a.items = (const uint4 *)pair_->items + 2 * (size_t)tab[l];
                a.n_items = tab[l + 1] - tab[l];
                a.disc_rows = pair_->disc_rows + tab[l + 2];
                a.n_disc = tab[l + 3] - tab[l + 2];
                a.val = V;
                a.vld = M;
                a.stride = stride_;
                a.pairs = pairs;
                a.pair_blocks = sh.pair_blocks;
                a.sem = net_.sem_flags;
                a.discard = opt_discard_ ? 1 : 0;
                a.hints = opt_l2_hints_ ? 1 : 0;
                a.pdl = opt_pdl_;
                if (a.n_items == 0) continue;
                cudaLaunchConfig_t cfg = {};
                // CTAs of 128 threads, like k_eval2 with two gates per thread
                dim3 blk = sh.block;
                if (blk.x <= 128) blk.y = 128 / blk.x;
                cfg.gridDim = dim3(sh.pair_blocks * ((a.n_items + blk.y - 1) / blk.y));
                cfg.blockDim = blk;
                cfg.stream = ST;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at;
                cfg.numAttrs = a.pdl ? 1 : 0;
                cudaLaunchKernelEx(&cfg, k_eval_pair, a);
                nodes++;
            }
        }
        for (const Level &lv : net_.levels) {
            if (use_pair) break;
            if (lv.g2_end > lv.g2_begin) {
                Eval2Args a{d_fan0_, d_fan1_, d_fan2_, d_kind2_, V, M, V, M, nullptr, lv.g2_begin, lv.g2_end, net_.n_src, stride_, pairs,
                            sh.pair_blocks, d_kv_, nullptr, nullptr, d_all_valid_};
                a.sem = net_.sem_flags;
                a.far = opt_l2_hints_ ? d_far2_ : nullptr;
                if (opt_discard_ && stride_ % 16 == 0) {
                    const size_t li = (size_t)(&lv - net_.levels.data());
                    if (li < disc_never_.size() && stride_ >= 16) {  // rows of this level that nobody reads (gate indices)
                        a.now_begin = disc_never_[li].begin - net_.n_src;
                        a.now_end = disc_never_[li].end - net_.n_src;
                    }
                    if (li >= 2 && li - 2 < disc_short_.size()) {  // rows two levels below whose only readers were the level below
                        a.disc_begin[1] = disc_short_[li - 2].begin;
                        a.disc_end[1] = disc_short_[li - 2].end;
                    }
                }
                // U gates per thread: 4 when the level still fills the machine several times over, else 2
                // (more, smaller CTAs: a level of a small tile is only a wave or two of work)
                const uint32_t n_g = lv.g2_end - lv.g2_begin;
                const uint32_t ctas4 = sh.pair_blocks * ((n_g + sh.block.y * 4 - 1) / (sh.block.y * 4));
                const uint32_t U = (opt_u_ ? opt_u_ : (ctas4 >= 148 * 3 * 4 ? 4u : 2u));
                // two gates per thread (the small-tile regime): CTAs of 128 threads — ten per SM instead of five of 256 — pack the
                // tails of three slots' overlapping launches better (+2 % on config 3; 64 threads: worse)
                dim3 blk = sh.block;
                if (U == 2 && blk.x <= 128) blk.y = 128 / blk.x;
                const uint32_t per = blk.y * U;
                const uint32_t gb = (n_g + per - 1) / per;
                a.pdl = opt_pdl_;
                auto launch = [&](void (*kern)(const Eval2Args)) {
                    if (!a.pdl) {
                        kern<<<sh.pair_blocks * gb, blk, 0, ST>>>(a);
                        return;
                    }
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(sh.pair_blocks * gb);
                    cfg.blockDim = blk;
                    cfg.stream = ST;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = 1;
                    cudaLaunchKernelEx(&cfg, kern, a);
                };
                if (U == 4) launch(kv ? k_eval2<4, false, true> : k_eval2<4, false, false>);
                else launch(kv ? k_eval2<2, false, true> : k_eval2<2, false, false>);
                nodes++;
            }
            if (lv.g3_end > lv.g3_begin) {
                Eval2Args a{d_fan0_, d_fan1_, d_fan2_, d_kind2_, V, M, V, M, nullptr, lv.g3_begin, lv.g3_end, net_.n_src, stride_, pairs,
                            sh.pair_blocks, d_kv_, nullptr, nullptr, d_all_valid_};
                a.sem = net_.sem_flags;
                const uint32_t per = sh.block.y * 2;
                const uint32_t gb = (lv.g3_end - lv.g3_begin + per - 1) / per;
                if (kv)
                    k_eval2<2, false, true, 3><<<sh.pair_blocks * gb, sh.block, 0, ST>>>(a);
                else
                    k_eval2<2, false, false, 3><<<sh.pair_blocks * gb, sh.block, 0, ST>>>(a);
                nodes++;
            }
            if (lv.wide_end > lv.wide_begin) {
                WideArgs a{d_wide_off_, d_wide_in_, d_wide_kind_, d_wide_nsel_, V, M, V, M, nullptr, lv.wide_begin, lv.wide_end,
                           net_.n_src + net_.n_fast(), stride_, pairs, sh.pair_blocks, kv ? d_kv_ : nullptr, nullptr, nullptr, d_all_valid_};
                a.aux = d_wide_aux_;
                a.sem = net_.sem_flags;
                const uint32_t gb = (lv.wide_end - lv.wide_begin + sh.block.y - 1) / sh.block.y;
                k_eval_wide<false><<<sh.pair_blocks * gb, sh.block, 0, ST>>>(a);
                nodes++;
            }
        }
        if (capturing) {
            cudaGraph_t graph = nullptr;
            CU_TRY(cudaStreamEndCapture(ST, &graph));
            cudaGraphExec_t exec = nullptr;
            cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) {
                set_last_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
                return B2_ERR_CUDA;
            }
            level_graph_exec_ = exec;
            level_graph_key_ = gkey;
            level_graph_nodes_ = nodes;
            CU_TRY(cudaGraphLaunch(exec, ST));
        } else {
            CU_TRY(cudaGetLastError());
        }
        g_launches.fetch_add(nodes, std::memory_order_relaxed);
    }
    cold_ = false;
    frontier_valid_ = false;
    last_mode_ = 1;
    last_apps_ = 1;
    elided_ = kv;  // a plain pass rewrites every valid plane, so nothing stays elided
    if (extra_rows_.empty()) {
        if (async) return B2_RUN_OK;  // no wire can conflict; status words are refreshed lazily
        CU_TRY(cudaMemsetAsync(d_conflict_, 0, (size_t)stride_ * 8, ST));
        CU_TRY(cudaMemsetAsync(d_unsettled_, 0, (size_t)stride_ * 8, ST));
        h_flags_->any_live = h_flags_->any_conflict = h_flags_->any_unsettled = 0;
        return B2_RUN_OK;
    }
    if ((rc0 = sync_extra_tables())) return rc0;
    CU_TRY(cudaMemsetAsync(d_conflict_, 0, (size_t)stride_ * 8, ST));
    k_clear_flags<<<1, 1, 0, ST>>>(d_flags_);
    LAUNCHED();
    k_extra_drive<<<(stride_ + 255) / 256, 256, 0, ST>>>(d_extra_rows_, (uint32_t)extra_rows_.size(), xdrv_val_, xdrv_vld_, V, M,
                                                         d_conflict_, stride_, 0, opt_lenient_ ? 1 : 0);
    LAUNCHED();
    k_levelised_status<<<(stride_ + 255) / 256, 256, 0, ST>>>(d_conflict_, d_unsettled_, d_flags_, stride_, n_lanes_);
    LAUNCHED();
    if ((rc0 = drives_end())) return rc0;
    CU_TRY(cudaMemcpyAsync(h_flags_, d_flags_, sizeof(RunFlags), cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaStreamSynchronize(ST));
    return h_flags_->any_conflict ? B2_RUN_CONFLICT : B2_RUN_OK;
}

// Continues the unit-delay loop of the current run from application k+1 on the device (k_unit_delay_loop).
// finished: the run ended there (result = its b2_run_result); otherwise the kernel handed back before an
// application whose work list is long, and the host loop goes on from the returned k.
int Engine::run_device_loop(uint64_t &k, uint64_t last, bool &finished, int &result) {
    const uint32_t pairs = stride_ / 2;
    const Shape sh = shape_for(pairs);
    const uint32_t n_res = net_.n_res(), res_base = net_.n_src + net_.n_fast() + net_.n_wide();
    CU_TRY(cudaMemsetAsync(d_rowchg_[chg_w_], 0, rowchg_bytes(), ST));
    LoopArgs la;
    la.fan0 = d_fan0_; la.fan1 = d_fan1_; la.fan2 = d_fan2_; la.kind2 = d_kind2_;
    la.wide = WideArgs{d_wide_off_, d_wide_in_, d_wide_kind_, d_wide_nsel_, nullptr, nullptr, nullptr, nullptr, d_changed_, 0,
                       net_.n_wide(), net_.n_src + net_.n_fast(), stride_, pairs, sh.pair_blocks, nullptr, nullptr, nullptr, nullptr};
    la.res = ResolveArgs{d_res_off_, d_res_in_, d_res_xslot_, xdrv_val_, xdrv_vld_, nullptr, nullptr, nullptr, nullptr, d_changed_,
                         d_conf_step_, nullptr, n_res, res_base, stride_, pairs, sh.pair_blocks, 0, opt_lenient_ ? 1 : 0};
    for (int i = 0; i < 2; i++) {
        la.val[i] = val_[i];
        la.vld[i] = vld_[i];
        la.chg[i] = d_rowchg_[i];
    }
    la.drv_val = drv_val_; la.drv_vld = drv_vld_;
    la.list2 = (uint4 *)d_act_list_[0]; la.list3 = (uint4 *)d_act_list_[1]; la.listc = d_act_list_[2];
    la.listw = d_act_list_[3];
    la.cnt = (uint32_t *)(d_loop_result_ + 1);  // behind the result block: 8 list counters, then the 8 period-2 words
    CU_TRY(cudaMemsetAsync(la.cnt, 0, 16 * sizeof(uint32_t), ST));
    la.p2 = opt_period2_ ? la.cnt + 8 : nullptr;
    la.changed = d_changed_; la.conf_step = d_conf_step_; la.live = d_live_; la.conflict = d_conflict_; la.unsettled = d_unsettled_;
    la.flags = d_flags_;
    la.result = d_loop_result_;
    la.n_src = net_.n_src; la.n_g2 = net_.n_g2(); la.n_fast = net_.n_fast(); la.n_wide = net_.n_wide(); la.n_res = n_res;
    la.stride = stride_; la.pairs = pairs; la.pair_blocks = sh.pair_blocks; la.rowchg_bytes = (uint32_t)rowchg_bytes();
    la.n_lanes = n_lanes_; la.k0 = k; la.last = last;
    la.bail_row_words = 4 * kLoopEntryRowWords;
    la.sem = net_.sem_flags;
    la.wide.sem = net_.sem_flags;
    la.wide.aux = d_wide_aux_;
    la.cur = cur_; la.chg_w = chg_w_; la.frontier_valid = frontier_valid_ ? 1 : 0;
    void *params[] = {&la};
    const uint32_t grid = sh.pair_blocks * (loop_ctas_ / sh.pair_blocks);
    {
        // A refused cooperative launch (co-residency cannot be granted, e.g. under a restricted SM partition) is not an
        // error of the run: the host-driven loop does the same work, so this engine stays on it from now on.
        const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_unit_delay_loop, dim3(grid), sh.block, params, 0, ST);
        if (e != cudaSuccess) {
            cudaGetLastError();
            loop_ctas_ = 0;
            finished = false;
            return B2_OK;
        }
    }
    LAUNCHED();
    CU_TRY(cudaMemcpyAsync(h_loop_result_, d_loop_result_, sizeof(LoopResult), cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaStreamSynchronize(ST));
    const LoopResult *lr = (const LoopResult *)h_loop_result_;
    last_apps_ += lr->k - k - lr->skipped;
    k = lr->k;
    cur_ = lr->cur;
    chg_w_ = lr->chg_w;
    frontier_valid_ = true;
    finished = lr->bailed == 0;
    if (finished) {
        if (k == last) frontier_valid_ = false;  // ended on the check-only application: the other buffer is one step ahead
        h_flags_->any_conflict = lr->any_conflict;
        h_flags_->any_unsettled = lr->any_unsettled;
        result = lr->any_conflict ? B2_RUN_CONFLICT : (lr->any_unsettled ? B2_RUN_MAX_STEPS_REACHED : B2_RUN_OK);
    }
    return B2_OK;
}

int Engine::run_jacobi(uint64_t max_steps) {
    int rc = ensure_second_buffer();
    if (rc) return rc;
    if ((rc = materialise_valid())) return rc;
    if ((rc = drives_begin())) return rc;
    const uint32_t pairs = stride_ / 2;
    const Shape sh = shape_for(pairs);
    // step_sim calls gsim allows after begin_sim: max_steps + 1 with its `steps > max_steps` check
    const uint64_t N = opt_budget_inclusive_ ? max_steps : (max_steps == UINT64_MAX ? UINT64_MAX : max_steps + 1);
    const uint64_t last = N >= UINT64_MAX - 2 ? UINT64_MAX : N + 2;  // the final, check-only application
    const uint32_t n_slots = (uint32_t)extra_rows_.size(), n_extra = n_plain_extra_, n_res = net_.n_res();
    if ((rc = sync_extra_tables())) return rc;
    const uint32_t res_base = net_.n_src + net_.n_fast() + net_.n_wide();
    const uint32_t res_gb = n_res ? (n_res + sh.block.y - 1) / sh.block.y : 0;
    const uint32_t sgrid = (stride_ + 255) / 256;
    k_init_status<<<sgrid, 256, 0, ST>>>(d_changed_, d_conf_step_, d_live_, d_conflict_, d_unsettled_, d_flags_, stride_, n_lanes_, 0);
    LAUNCHED();
    last_mode_ = 2;
    last_apps_ = 0;
    h_flags_->any_live = h_flags_->any_conflict = h_flags_->any_unsettled = 0;
    uint64_t k = 0;
    uint32_t newest = 0;
    if (cold_) {
        // W_1 = resolve(D, all outputs High-Z) = D
        uint64_t *V = val_[cur_], *M = vld_[cur_];
        const size_t row_words = (size_t)net_.n_rows * stride_;
        // (A.6-5 alternative: component outputs start Undefined instead of High-Z, so every driven wire is X in W_1 and a drive
        //  on it — or a second driver — conflicts right there; the conflict words are counted by the first bookkeeping step)
        const bool init_x = (net_.sem_flags & SEM_INIT_X) != 0;
        CU_TRY(cudaMemsetAsync(V, init_x ? 0xFF : 0, row_words * 8, ST));
        CU_TRY(cudaMemsetAsync(M, 0, row_words * 8, ST));
        if (net_.n_src) {
            const uint32_t rb = (net_.n_src + sh.block.y - 1) / sh.block.y;
            k_apply_drive<false><<<sh.pair_blocks * rb, sh.block, 0, ST>>>(drv_val_, drv_vld_, nullptr, nullptr, V, M, nullptr,
                                                                            net_.n_src, stride_, pairs, sh.pair_blocks);
            LAUNCHED();
        }
        if (n_extra) {
            k_extra_drive<<<sgrid, 256, 0, ST>>>(d_extra_rows_, n_slots, xdrv_val_, xdrv_vld_, V, M, d_conf_step_, stride_, init_x ? 0 : 1,
                                                 opt_lenient_ ? 1 : 0);
            LAUNCHED();
        }
        if (n_res) {
            ResolveArgs ra{d_res_off_, d_res_in_, d_res_xslot_, xdrv_val_, xdrv_vld_, V, M, V, M, nullptr, d_conf_step_, nullptr,
                           n_res, res_base, stride_, pairs, sh.pair_blocks, init_x ? 0 : 1, opt_lenient_ ? 1 : 0};
            k_resolve<<<sh.pair_blocks * res_gb, sh.block, 0, ST>>>(ra);
            LAUNCHED();
        }
        cold_ = false;
        frontier_valid_ = false;
        k = 1;
        if (net_.n_fast() + net_.n_wide() == 0) {
            if ((rc = drives_end())) return rc;
            return B2_RUN_OK;
        }
    }
    // Frontier (event-driven at row granularity): a gate is re-evaluated only if one of its fan-in rows, or
    // its own output row, changed in the previous application.  The flags persist across runs — a settled
    // run leaves them all zero, so the next run starts from just the source rows whose drive changed.
    // Off when gate outputs carry external drives (their merge bypasses the flags).
    // (work-list entries pack the gate index into 28 bits next to its kind)
    const bool use_frontier = opt_frontier_ && n_extra == 0 && net_.n_fast() < (1u << 28);
    if (use_frontier && !d_rowchg_[0]) {
        // each flag buffer carries the two list counters of its application behind the flags: one memset clears both
        CU_TRY(cudaMalloc((void **)&d_rowchg_[0], rowchg_bytes()));
        CU_TRY(cudaMalloc((void **)&d_rowchg_[1], rowchg_bytes()));
        CU_TRY(cudaMalloc((void **)&d_act_list_[0], (size_t)std::max<uint32_t>(net_.n_g2(), 1) * 16));
        CU_TRY(cudaMalloc((void **)&d_act_list_[1], (size_t)std::max<uint32_t>(net_.n_g3(), 1) * 16));
        CU_TRY(cudaMalloc((void **)&d_act_list_[2], (size_t)std::max<uint32_t>(net_.n_fast(), 1) * 4));
        CU_TRY(cudaMalloc((void **)&d_act_list_[3], (size_t)std::max<uint32_t>(net_.n_wide(), 1) * 4));
        store_bytes_ += 2 * rowchg_bytes() + (size_t)net_.n_fast() * 20 + (size_t)net_.n_wide() * 4;
        frontier_valid_ = false;
    }
    if (!use_frontier) frontier_valid_ = false;
    // Light applications (a short work list times the row length) run in the device-side loop, heavy ones as
    // separate fully occupied kernels; the run moves between the two as the activity changes.
    bool device_ok = opt_device_loop_ && use_frontier && loop_ctas_ >= sh.pair_blocks;
    bool want_device = device_ok && (uint64_t)(net_.n_fast() + net_.n_wide()) * stride_ <= kLoopEntryRowWords;
    const uint64_t first_k = k + 1;
    uint64_t lag_base = first_k;  // first application of the current host-driven stretch
    uint32_t light_streak = 0;
    for (;;) {
        if (want_device) {
            bool finished = false;
            int result = B2_RUN_OK;
            if ((rc = run_device_loop(k, last, finished, result))) return rc;
            if (finished) {
                if ((rc = drives_end())) return rc;
                return result;
            }
            want_device = false;
            light_streak = 0;
            lag_base = k + 1;
            if (loop_ctas_ == 0) device_ok = false;  // the cooperative launch was refused
        }
        k++;
        last_apps_++;
        const bool is_last = k == last;
        const uint64_t *AV = val_[cur_], *AM = vld_[cur_];
        uint64_t *BV = val_[cur_ ^ 1], *BM = vld_[cur_ ^ 1];
        // changed-row frontier: this application reads the flags of the previous one and writes its own
        uint8_t *chg_cur = use_frontier ? d_rowchg_[chg_w_] : nullptr;
        const uint8_t *chg_prev = (use_frontier && frontier_valid_) ? d_rowchg_[chg_w_ ^ 1] : nullptr;
        if (use_frontier) CU_TRY(cudaMemsetAsync(chg_cur, 0, rowchg_bytes(), ST));
        uint32_t *act_counts = use_frontier ? (uint32_t *)(chg_cur + rowchg_bytes() - 16) : nullptr;
        // Drives cannot change inside a run: after two applications both buffers hold them on the source rows.
        if (net_.n_src && k < first_k + 2) {
            const uint32_t rb = (net_.n_src + sh.block.y - 1) / sh.block.y;
            k_apply_drive<true><<<sh.pair_blocks * rb, sh.block, 0, ST>>>(drv_val_, drv_vld_, AV, AM, BV, BM, d_changed_, net_.n_src,
                                                                           stride_, pairs, sh.pair_blocks, nullptr, 0, chg_cur);
            LAUNCHED();
        }
        if (chg_prev && net_.n_fast()) {
            // event-driven: compact the gates whose fan-in (or own row) changed, evaluate only those
            k_frontier_scan<<<(net_.n_fast() + 255) / 256, 256, 0, ST>>>(d_fan0_, d_fan1_, d_fan2_, net_.n_g2(), net_.n_fast(), net_.n_src,
                                                                         d_kind2_, chg_prev, (uint4 *)d_act_list_[0], (uint4 *)d_act_list_[1], d_act_list_[2],
                                                                         act_counts);
            LAUNCHED();
            const uint32_t per = sh.block.y * 2;
            // persistent grids: exactly as many CTAs as are resident at once (a partial second wave would double the time)
            auto slots = [&](int ctas_per_sm) { return std::max<uint32_t>(1, (uint32_t)(sm_count_ * ctas_per_sm) / sh.pair_blocks); };
            uint32_t gbs_max = slots(occ_list_[0]);
            find_hot_rows();
            auto set_hot = [&](Eval2Args &a) {
                a.hot_n = opt_hot_ ? n_hot_ : 0;
                for (uint32_t h = 0; h < a.hot_n; h++) a.hot_row[h] = hot_rows_[h];
            };
            if (net_.n_g2()) {
                Eval2Args a{d_fan0_, d_fan1_, d_fan2_, d_kind2_, AV, AM, BV, BM, d_changed_, 0, net_.n_g2(), net_.n_src, stride_, pairs,
                            sh.pair_blocks, nullptr, nullptr, chg_cur, nullptr};
                a.sem = net_.sem_flags;
                set_hot(a);
                const uint32_t gbs = std::min<uint32_t>((net_.n_g2() + per - 1) / per, gbs_max);
                if (a.hot_n) k_eval_list_hot<2><<<sh.pair_blocks * gbs, sh.block, 0, ST>>>(a, (const uint4 *)d_act_list_[0], act_counts);
                else k_eval_list<2><<<sh.pair_blocks * gbs, sh.block, 0, ST>>>(a, (const uint4 *)d_act_list_[0], act_counts);
                LAUNCHED();
            }
            gbs_max = slots(occ_list_[1]);
            if (net_.n_g3()) {
                Eval2Args a{d_fan0_, d_fan1_, d_fan2_, d_kind2_, AV, AM, BV, BM, d_changed_, net_.n_g2(), net_.n_fast(), net_.n_src, stride_,
                            pairs, sh.pair_blocks, nullptr, nullptr, chg_cur, nullptr};
                a.sem = net_.sem_flags;
                set_hot(a);
                const uint32_t gbs = std::min<uint32_t>((net_.n_g3() + per - 1) / per, gbs_max);
                if (a.hot_n) k_eval_list_hot<3><<<sh.pair_blocks * gbs, sh.block, 0, ST>>>(a, (const uint4 *)d_act_list_[1], act_counts + 1);
                else k_eval_list<3><<<sh.pair_blocks * gbs, sh.block, 0, ST>>>(a, (const uint4 *)d_act_list_[1], act_counts + 1);
                LAUNCHED();
            }
            gbs_max = slots(occ_list_[2]);
            {
                const uint32_t gbs = std::min<uint32_t>((net_.n_fast() + sh.block.y - 1) / sh.block.y, gbs_max);
                k_copy_list<<<sh.pair_blocks * gbs, sh.block, 0, ST>>>(d_act_list_[2], act_counts + 2, AV, AM, BV, BM, stride_, pairs,
                                                                       sh.pair_blocks);
                LAUNCHED();
            }
        } else {
            if (net_.n_g2()) {
                Eval2Args a{d_fan0_, d_fan1_, d_fan2_, d_kind2_, AV, AM, BV, BM, d_changed_, 0, net_.n_g2(), net_.n_src, stride_, pairs,
                            sh.pair_blocks, nullptr, chg_prev, chg_cur, nullptr};
                a.sem = net_.sem_flags;
                const uint32_t per = sh.block.y * EVAL2_U;
                const uint32_t gb = (net_.n_g2() + per - 1) / per;
                k_eval2<EVAL2_U, true, false><<<sh.pair_blocks * gb, sh.block, 0, ST>>>(a);
                LAUNCHED();
            }
            if (net_.n_g3()) {
                Eval2Args a{d_fan0_, d_fan1_, d_fan2_, d_kind2_, AV, AM, BV, BM, d_changed_, net_.n_g2(), net_.n_fast(), net_.n_src, stride_,
                            pairs, sh.pair_blocks, nullptr, chg_prev, chg_cur, nullptr};
                a.sem = net_.sem_flags;
                const uint32_t per = sh.block.y * 2;
                const uint32_t gb = (net_.n_g3() + per - 1) / per;
                k_eval2<2, true, false, 3><<<sh.pair_blocks * gb, sh.block, 0, ST>>>(a);
                LAUNCHED();
            }
        }
        if (net_.n_wide()) {
            WideArgs a{d_wide_off_, d_wide_in_, d_wide_kind_, d_wide_nsel_, AV, AM, BV, BM, d_changed_, 0, net_.n_wide(),
                       net_.n_src + net_.n_fast(), stride_, pairs, sh.pair_blocks, nullptr, chg_prev, chg_cur, nullptr};
            a.aux = d_wide_aux_;
            a.sem = net_.sem_flags;
            if (chg_prev) {  // event-driven: list the wide gates with a changed input, evaluate the list
                k_frontier_scan_wide<<<(net_.n_wide() + 255) / 256, 256, 0, ST>>>(d_wide_off_, d_wide_in_, net_.n_wide(),
                                                                                  net_.n_src + net_.n_fast(), chg_prev, d_act_list_[3],
                                                                                  act_counts + 3);
                LAUNCHED();
                const uint32_t slots_w = std::max<uint32_t>(1, (uint32_t)(sm_count_ * occ_list_[3]) / sh.pair_blocks);
                const uint32_t gbs = std::min<uint32_t>((net_.n_wide() + sh.block.y - 1) / sh.block.y, slots_w);
                k_eval_wide_list<<<sh.pair_blocks * gbs, sh.block, 0, ST>>>(a, d_act_list_[3], act_counts + 3);
                LAUNCHED();
            } else {
                const uint32_t gb = (net_.n_wide() + sh.block.y - 1) / sh.block.y;
                k_eval_wide<true><<<sh.pair_blocks * gb, sh.block, 0, ST>>>(a);
                LAUNCHED();
            }
        }
        if (use_frontier) {
            frontier_valid_ = true;
            chg_w_ ^= 1;
        }
        if (n_res) {
            ResolveArgs ra{d_res_off_, d_res_in_, d_res_xslot_, xdrv_val_, xdrv_vld_, AV, AM, BV, BM, d_changed_, d_conf_step_, chg_cur,
                           n_res, res_base, stride_, pairs, sh.pair_blocks, 0, opt_lenient_ ? 1 : 0};
            k_resolve<<<sh.pair_blocks * res_gb, sh.block, 0, ST>>>(ra);
            LAUNCHED();
        }
        if (n_extra) {
            k_extra_drive<<<sgrid, 256, 0, ST>>>(d_extra_rows_, n_slots, xdrv_val_, xdrv_vld_, BV, BM, d_conf_step_, stride_, 0,
                                                 opt_lenient_ ? 1 : 0);
            LAUNCHED();
        }
        // The lane bookkeeping of application k goes to flag slot k % kFlagSlots; the host reads it
        // kFlagLag applications later, so the device never waits for the host.  Applications queued
        // past the point where every lane has settled are no-ops on a fixed point.
        const uint32_t slot = (uint32_t)(k % kFlagSlots);
        k_status_step<<<sgrid, 256, 0, ST>>>(d_changed_, d_conf_step_, d_live_, d_conflict_, d_unsettled_, d_flags_ + slot, stride_,
                                             n_lanes_, is_last ? 0 : 1, is_last ? 1 : 0, d_flags_ + (k + 1) % kFlagSlots,
                                             chg_prev ? act_counts : nullptr);
        LAUNCHED();
        CU_TRY(cudaMemcpyAsync(h_flags_ + slot, d_flags_ + slot, sizeof(RunFlags), cudaMemcpyDeviceToHost, ST));
        CU_TRY(cudaEventRecord(EV(ev_flags_[slot]), ST));
        if (!is_last) cur_ ^= 1;  // check-only last application: MaxStepsReached lanes keep W_{N+1}
        newest = slot;
        bool stop = is_last;
        if (!stop && k >= lag_base + kFlagLag) {
            const uint32_t old = (uint32_t)((k - kFlagLag) % kFlagSlots);
            CU_TRY(cudaEventSynchronize(EV(ev_flags_[old])));
            if (!h_flags_[old].any_live) stop = true;
            else if (device_ok && (uint64_t)h_flags_[old].active * stride_ <= kLoopEntryRowWords) {
                // a long tail of light applications (oscillating lanes): worth a cooperative launch; a settle
                // that is about to end is not
                if (++light_streak >= 4) want_device = true;
            } else {
                light_streak = 0;
            }
        }
        if (stop) break;
    }
    CU_TRY(cudaStreamSynchronize(ST));
    if (k == last) frontier_valid_ = false;  // ended on the check-only application: the other buffer is one step ahead
    const RunFlags fin = h_flags_[newest];  // conflict flags are cumulative; unsettled exists only on the last application
    if ((rc = drives_end())) return rc;
    if (fin.any_conflict) return B2_RUN_CONFLICT;
    if (fin.any_unsettled) return B2_RUN_MAX_STEPS_REACHED;
    return B2_RUN_OK;
}

// Shared-memory budget of the resident kernel (bytes of dynamic shared memory per CTA).
static const size_t kResidentSmem = 200 * 1024;

// 0 = not applicable, else lane-words per CTA
uint32_t Engine::resident_wpc() const {
    if (net_.n_rows == 0) return 0;
    const size_t per_word = (size_t)net_.n_rows * 32 + 5 * 8;
    if (per_word > kResidentSmem) return 0;
    size_t wpc = kResidentSmem / per_word;
    const size_t spread = ((size_t)stride_ + 147) / 148;  // enough CTAs to cover the SMs first
    if (stride_ <= 2 && wpc >= stride_) return stride_;   // one lane-word (+ its padding word): one CTA, which also packs the snapshot
    if (wpc > spread) wpc = spread;
    if (wpc > 64) wpc = 64;
    return (uint32_t)(wpc ? wpc : 1);
}

// Page-locked host memory the device can store the run flags and the lane-0 snapshot into (one-CTA resident runs).
bool Engine::direct_ready() {
    if (!opt_direct_ || !h_flags_dev_) return false;
    const size_t need = 2 * (((size_t)net_.n_rows / 8 + 1 + 7) & ~(size_t)7) + 64;
    if (h_snap_cap_ >= need) return true;
    if (cudaStreamSynchronize(ST) != cudaSuccess) return false;
    if (h_snap_) cudaFreeHost(h_snap_);
    h_snap_ = h_snap_dev_ = nullptr;
    h_snap_cap_ = 0;
    if (cudaHostAlloc((void **)&h_snap_, need, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&h_snap_dev_, h_snap_, 0) != cudaSuccess) {
        cudaGetLastError();
        if (h_snap_) cudaFreeHost(h_snap_);
        h_snap_ = h_snap_dev_ = nullptr;
        return false;
    }
    h_snap_cap_ = need;
    return true;
}

int Engine::run_resident(uint64_t max_steps, bool levelised, const void *inline_bits, uint32_t n_inline) {
    const uint32_t wpc = resident_wpc();
    // step_sim calls gsim allows after begin_sim: max_steps + 1 with its `steps > max_steps` check
    const uint64_t N = opt_budget_inclusive_ ? max_steps : (max_steps == UINT64_MAX ? UINT64_MAX : max_steps + 1);
    const uint64_t last = N >= UINT64_MAX - 2 ? UINT64_MAX : N + 2;
    const uint32_t n_extra = (uint32_t)extra_rows_.size();
    int rc;
    if ((rc = materialise_valid())) return rc;
    if ((rc = drives_begin())) return rc;
    if ((rc = sync_extra_tables())) return rc;
    const uint32_t ctas = (stride_ + wpc - 1) / wpc;
    const bool with_snapshot = T_ <= 64;  // interactive-sized runs: get_wire_state will want lane 0 next
    // one CTA and page-locked memory the device can store into: flags and snapshot go straight to the host
    const size_t snap_bytes = (size_t)net_.n_rows / 8 + 1;
    const bool direct = ctas == 1 && with_snapshot && direct_ready();
    if (n_inline && !direct) return B2_ERR_INVALID_ARG;  // Engine::run checked the same conditions
    // flag slot 0 takes the run's flags, slot 1 the application count: one clear before, one copy after the kernel
    if (!direct) CU_TRY(cudaMemsetAsync(d_flags_, 0, 2 * sizeof(RunFlags), ST));
    ResidentArgs a;
    a.host_flags = direct ? (RunFlags *)h_flags_dev_ : nullptr;
    a.n_inline = n_inline;
    for (uint32_t i = 0; i < n_inline; i++) a.inl[i] = ((const PendingBit *)inline_bits)[i];
    a.fan0 = d_fan0_; a.fan1 = d_fan1_; a.fan2 = d_fan2_; a.kind2 = d_kind2_;
    a.woff = d_wide_off_; a.win = d_wide_in_; a.wkind = d_wide_kind_; a.wnsel = d_wide_nsel_; a.waux = d_wide_aux_;
    a.xrows = d_extra_rows_; a.xval = xdrv_val_; a.xvld = xdrv_vld_;
    a.res_off = d_res_off_; a.res_in = d_res_in_; a.res_xslot = d_res_xslot_; a.n_res = net_.n_res();
    a.dval = drv_val_; a.dvld = drv_vld_;
    a.val = val_[cur_]; a.vld = vld_[cur_];
    a.conflict = d_conflict_; a.unsettled = d_unsettled_;
    a.flags = d_flags_;
    a.max_apps = (unsigned long long *)(d_flags_ + 1);
    a.n_src = net_.n_src; a.n_g2 = net_.n_g2(); a.n_g3 = net_.n_g3(); a.n_wide = net_.n_wide(); a.n_extra = n_extra; a.n_rows = net_.n_rows;
    a.stride = stride_; a.wpc = wpc;
    a.n_lanes = n_lanes_;
    a.last = last;
    a.cold = cold_ ? 1 : 0;
    a.levels = d_levels_;
    a.n_levels = (uint32_t)net_.levels.size();
    a.levelised = levelised ? 1 : 0;
    a.lenient = opt_lenient_ ? 1 : 0;
    a.sem = net_.sem_flags;
    a.period2 = opt_period2_ ? 1 : 0;
    const size_t smem = ((size_t)net_.n_rows * wpc * 4 + (size_t)wpc * 5) * 8;
    if (!resident_attr_set_) {
        CU_TRY(cudaFuncSetAttribute(k_resident_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kResidentSmem + 4096)));
        resident_attr_set_ = true;
    }
    const size_t items = (size_t)net_.n_rows * wpc;
    const uint32_t threads = items >= 1024 ? 1024 : (items >= 256 ? 512 : 128);
    a.snap0 = a.snap1 = nullptr;
    a.snap_bytes = 0;
    const size_t snap_pitch = (snap_bytes + 7) & ~(size_t)7;  // the direct path stores 8 bytes at a time
    if (direct) {
        a.snap0 = h_snap_dev_;
        a.snap1 = h_snap_dev_ + snap_pitch;
        a.snap_bytes = (uint32_t)snap_bytes;
        snap_.resize(2 * snap_bytes);
    } else if (with_snapshot) {
        if (pack_cap_ < 2 * snap_bytes) {
            CU_TRY(cudaStreamSynchronize(ST));
            cudaFree(d_pack_);
            d_pack_ = nullptr;
            pack_cap_ = 0;
            CU_TRY(cudaMalloc((void **)&d_pack_, 2 * snap_bytes + 64));
            pack_cap_ = 2 * snap_bytes + 64;
        }
        a.snap0 = d_pack_;
        a.snap1 = d_pack_ + snap_bytes;
        a.snap_bytes = (uint32_t)snap_bytes;
        snap_.resize(2 * snap_bytes);
    }
    k_resident_run<<<ctas, threads, smem, ST>>>(a);
    LAUNCHED();
    if (!direct) {
        CU_TRY(cudaMemcpyAsync(h_flags_, d_flags_, 2 * sizeof(RunFlags), cudaMemcpyDeviceToHost, ST));
        if (with_snapshot) CU_TRY(cudaMemcpyAsync(snap_.data(), d_pack_, 2 * snap_bytes, cudaMemcpyDeviceToHost, ST));
    }
    if ((rc = drives_end())) return rc;
    CU_TRY(cudaStreamSynchronize(ST));
    if (direct) {
        std::memcpy(snap_.data(), h_snap_, snap_bytes);
        std::memcpy(snap_.data() + snap_bytes, h_snap_ + snap_pitch, snap_bytes);
    }
    snap_valid_ = with_snapshot;
    cold_ = false;
    frontier_valid_ = false;
    last_mode_ = 3;
    last_apps_ = *(const unsigned long long *)(h_flags_ + 1);
    if (h_flags_->any_conflict) return B2_RUN_CONFLICT;
    if (h_flags_->any_unsettled) return B2_RUN_MAX_STEPS_REACHED;
    return B2_RUN_OK;
}

int Engine::run(uint64_t max_steps, uint64_t *conflict, uint64_t *unsettled, bool async) {
    int rc = ensure_lanes();
    if (rc) return -rc;
    // the one-lane server cycle: a few queued drives, a netlist that runs in one resident CTA — the drives ride in the kernel's
    // arguments instead of an upload and a launch of their own
    PendingBit inl[kInlineBits];
    uint32_t n_inline = 0;
    {
        const bool dag = !net_.cyclic && max_steps >= (uint64_t)net_.depth + 1;
        const bool resident = !net_.multi_driver && use_resident_ && resident_wpc() && !(dag && !(stride_ <= 4)) && T_ <= 64 &&
                              (stride_ + resident_wpc() - 1) / resident_wpc() == 1 && direct_ready();
        if (resident && !pending_.empty() && (rc = pending_to_bits(inl, kInlineBits, &n_inline))) return -rc;
    }
    if (n_inline == 0 && (rc = flush_pending_drives())) return -rc;
    if (side_pending_) {  // a gather of the previous run's rows may still be reading the store on the side stream
        if (cudaStreamWaitEvent(ST, EV(ev_side_done_), 0) != cudaSuccess) return -B2_ERR_CUDA;
        side_pending_ = false;
    }
    snap_valid_ = false;
    // acyclic, single-driver, and the step budget covers the depth: the unique fixed point, level by level
    const bool dag_ok = !net_.cyclic && max_steps >= (uint64_t)net_.depth + 1;
    int result;
    if (net_.multi_driver) {
        // two gate outputs on one wire are never High-Z: every lane conflicts on its first step
        k_init_status<<<(stride_ + 255) / 256, 256, 0, ST>>>(d_changed_, d_conf_step_, d_live_, d_conflict_, d_unsettled_, d_flags_,
                                                             stride_, n_lanes_, 1);
        g_launches.fetch_add(1);
        if (cudaGetLastError() != cudaSuccess) return -B2_ERR_CUDA;
        cold_ = false;
        last_mode_ = 0;
        last_apps_ = 0;
        result = B2_RUN_CONFLICT;
    } else if (dag_ok && !(stride_ <= 4 && resident_wpc() && use_resident_)) {
        result = run_levelised(async && !conflict && !unsettled);
    } else if (use_resident_ && resident_wpc()) {
        result = run_resident(max_steps, dag_ok, inl, n_inline);  // small netlist: the whole run_sim loop (or level sweep) on chip
    } else {
        result = run_jacobi(max_steps);
    }
    if (result < 0) return result;
    if (result > 2) return -result;  // a b2_status from CU_TRY
    if (conflict || unsettled) {
        if (conflict) {
            cudaError_t e = cudaMemcpyAsync(conflict, d_conflict_, (size_t)T_ * 8, cudaMemcpyDeviceToHost, ST);
            if (e != cudaSuccess) return -B2_ERR_CUDA;
        }
        if (unsettled) {
            cudaError_t e = cudaMemcpyAsync(unsettled, d_unsettled_, (size_t)T_ * 8, cudaMemcpyDeviceToHost, ST);
            if (e != cudaSuccess) return -B2_ERR_CUDA;
        }
        if (cudaStreamSynchronize(ST) != cudaSuccess) return -B2_ERR_CUDA;
    }
    return result;
}

// ---------------------------------------------------------------------------- Engine: read-back

int Engine::get_wire_lanes(uint32_t wire, uint32_t bit, uint32_t w0, uint32_t nw, uint64_t *value, uint64_t *valid) {
    if (wire >= net_.n_wires() || bit >= net_.width[wire]) return B2_ERR_INVALID_WIRE_ID;
    if (!value || !valid) return B2_ERR_NULL;
    int rc = ensure_lanes();
    if (rc) return rc;
    if ((uint64_t)w0 + nw > T_) return B2_ERR_RANGE;
    if ((rc = materialise_valid())) return rc;
    const size_t o = (size_t)net_.row_of_bit[net_.bit_off[wire] + bit] * stride_ + w0;
    CU_TRY(cudaMemcpyAsync(value, val_[cur_] + o, (size_t)nw * 8, cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaMemcpyAsync(valid, vld_[cur_] + o, (size_t)nw * 8, cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaStreamSynchronize(ST));
    return B2_OK;
}

int Engine::get_wire_state(uint32_t wire, uint32_t *p0, uint32_t *p1, size_t words) {
    if (wire >= net_.n_wires()) return B2_ERR_INVALID_WIRE_ID;
    if (!p0 || !p1) return B2_ERR_NULL;
    const uint32_t width = net_.width[wire], need = (width + 31) / 32;
    if (words < need) return B2_ERR_RANGE;
    if (!snap_valid_) {  // (the common case — every net read back after one eval — touches no CUDA call at all)
        int rc = ensure_lanes();
        if (rc) return rc;
        if ((rc = refresh_snapshot())) return rc;
    }
    for (uint32_t i = 0; i < need; i++) p0[i] = p1[i] = 0;
    const uint8_t *s0 = snap_.data(), *s1 = snap_.data() + snap_.size() / 2;
    for (uint32_t j = 0; j < width; j++) {
        const uint32_t row = net_.row_of_bit[net_.bit_off[wire] + j];
        p0[j >> 5] |= (uint32_t)((s0[row >> 3] >> (row & 7)) & 1) << (j & 31);
        p1[j >> 5] |= (uint32_t)((s1[row >> 3] >> (row & 7)) & 1) << (j & 31);
    }
    return B2_OK;
}

int Engine::gather_wires(const uint32_t *wires, uint32_t n, uint64_t *value, uint64_t *valid, size_t dst_stride, bool dst_is_device,
                         bool async) {
    if (n && (!wires || !value || !valid)) return B2_ERR_NULL;
    int rc = ensure_lanes();
    if (rc) return rc;
    if (n == 0) return B2_OK;
    if (dst_stride < T_) return B2_ERR_RANGE;
    if ((rc = upload_rows(wires, n, false))) return rc;
    auto launch = [&](uint64_t *dv, uint64_t *dm, size_t dstride) {
        const uint8_t *kvp = elided_ ? d_kv_ : nullptr;
        const bool vec = T_ % 2 == 0 && dstride % 2 == 0 && (((uintptr_t)dv | (uintptr_t)dm) & 15) == 0;
        if (vec) {
            dim3 block(64, 4), grid((T_ / 2 + 63) / 64, (n + 3) / 4);
            k_gather_rows<true><<<grid, block, 0, ST>>>(d_rows_, n, val_[cur_], vld_[cur_], dv, dm, stride_, 0, T_, dstride, kvp, d_all_valid_);
        } else {
            dim3 block(64, 4), grid((T_ + 63) / 64, (n + 3) / 4);
            k_gather_rows<false><<<grid, block, 0, ST>>>(d_rows_, n, val_[cur_], vld_[cur_], dv, dm, stride_, 0, T_, dstride, kvp, d_all_valid_);
        }
    };
    if (dst_is_device) {
        launch(value, valid, dst_stride);
        LAUNCHED();
        return B2_OK;
    }
    if ((rc = ensure_stage(true, (size_t)n * T_))) return rc;
    // main stream: gather into the out-stage once the previous D2H has drained it;
    // copy-out stream: D2H while the main stream goes on with the next tile
    CU_TRY(cudaStreamWaitEvent(ST, EV(ev_out_free_), 0));
    launch(stage_val_[1], stage_vld_[1], T_);
    LAUNCHED();
    CU_TRY(cudaEventRecord(EV(ev_gathered_), ST));
    CU_TRY(cudaStreamWaitEvent(COUT, EV(ev_gathered_), 0));
    CU_TRY(cudaMemcpy2DAsync(value, dst_stride * 8, stage_val_[1], (size_t)T_ * 8, (size_t)T_ * 8, n, cudaMemcpyDeviceToHost, COUT));
    CU_TRY(cudaMemcpy2DAsync(valid, dst_stride * 8, stage_vld_[1], (size_t)T_ * 8, (size_t)T_ * 8, n, cudaMemcpyDeviceToHost, COUT));
    CU_TRY(cudaEventRecord(EV(ev_out_free_), COUT));
    if (!async) CU_TRY(cudaStreamSynchronize(COUT));
    return B2_OK;
}

void Engine::set_discard(bool on, const uint32_t *keep_wires, uint32_t n) {
    opt_discard_ = on && !net_.cyclic;
    level_graph_key_ = 0;
    disc_short_.clear();
    disc_never_.clear();
    pair_.reset();
    pair_tried_ = false;
    keep_rows_.clear();
    if (!opt_discard_) return;
    std::vector<uint32_t> keep;
    for (uint32_t i = 0; i < n; i++)
        if (keep_wires[i] < net_.n_wires())
            for (uint32_t j = 0; j < net_.width[keep_wires[i]]; j++) keep.push_back(net_.row_of_bit[net_.bit_off[keep_wires[i]] + j]);
    std::sort(keep.begin(), keep.end());
    keep.erase(std::unique(keep.begin(), keep.end()), keep.end());
    keep_rows_ = keep;
    // the largest stretch of [b, e) without a kept row
    auto largest_gap = [&](uint32_t b, uint32_t e) -> RowRange {
        RowRange best{b, b};
        uint32_t start = b;
        auto it = std::lower_bound(keep.begin(), keep.end(), b);
        for (; it != keep.end() && *it < e; ++it) {
            if (*it - start > best.end - best.begin) best = RowRange{start, *it};
            start = *it + 1;
        }
        if (e - start > best.end - best.begin) best = RowRange{start, e};
        return best;
    };
    for (const Level &lv : net_.levels) {
        disc_short_.push_back(largest_gap(net_.n_src + lv.g2_short, net_.n_src + lv.g2_never));
        disc_never_.push_back(largest_gap(net_.n_src + lv.g2_never, net_.n_src + lv.g2_end));
    }
}

int Engine::gather_wires_multi(const uint32_t *wires, uint32_t n, const GatherDst *dsts, uint32_t n_dst) {
    if (n && (!wires || !dsts)) return B2_ERR_NULL;
    int rc = ensure_lanes();
    if (rc) return rc;
    if (n == 0 || n_dst == 0) return B2_OK;
    if ((rc = upload_rows(wires, n, false))) return rc;
    if (!side_stream_) {
        cudaStream_t st;
        CU_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        side_stream_ = st;
        for (void **e : {&ev_settled_, &ev_side_done_}) {
            cudaEvent_t ev;
            CU_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            *e = ev;
        }
    }
    cudaStream_t side = (cudaStream_t)side_stream_;
    CU_TRY(cudaEventRecord(EV(ev_settled_), ST));
    CU_TRY(cudaStreamWaitEvent(side, EV(ev_settled_), 0));
    const uint8_t *kvp = elided_ ? d_kv_ : nullptr;
    for (uint32_t first = 0; first < n_dst; first += kMaxGatherDst) {
        GatherDsts d;
        d.n = std::min<uint32_t>(kMaxGatherDst, n_dst - first);
        bool vec = T_ % 2 == 0;
        for (uint32_t k = 0; k < d.n; k++) {
            d.val[k] = dsts[first + k].val;
            d.vld[k] = dsts[first + k].vld;
            d.stride[k] = dsts[first + k].stride;
            if (d.stride[k] < T_) return B2_ERR_RANGE;
            vec = vec && d.stride[k] % 2 == 0 && (((uintptr_t)d.val[k] | (uintptr_t)d.vld[k]) & 15) == 0;
        }
        if (vec) {
            dim3 block(64, 4), grid((T_ / 2 + 63) / 64, (n + 3) / 4);
            k_gather_rows_multi<true><<<grid, block, 0, side>>>(d_rows_, n, val_[cur_], vld_[cur_], stride_, T_, d, kvp, d_all_valid_);
        } else {
            dim3 block(64, 4), grid((T_ + 63) / 64, (n + 3) / 4);
            k_gather_rows_multi<false><<<grid, block, 0, side>>>(d_rows_, n, val_[cur_], vld_[cur_], stride_, T_, d, kvp, d_all_valid_);
        }
        LAUNCHED();
    }
    CU_TRY(cudaEventRecord(EV(ev_side_done_), side));
    side_pending_ = true;
    return B2_OK;
}

// The main stream waits for the last asynchronous read-backs (gather_wires(..., async), gather_wires_multi) of this simulator.
int Engine::join_copies() {
    if (!device_ready_) return B2_OK;
    CU_TRY(cudaStreamWaitEvent(ST, EV(ev_out_free_), 0));
    if (side_pending_) {
        CU_TRY(cudaStreamWaitEvent(ST, EV(ev_side_done_), 0));
        side_pending_ = false;
    }
    return B2_OK;
}

int Engine::pack_sim_state(const uint32_t *wires, uint32_t n, uint8_t *p0, uint8_t *p1, size_t cap, uint64_t *bit_len) {
    if ((n && !wires) || !p0 || !p1) return B2_ERR_NULL;
    int rc = ensure_lanes();
    if (rc) return rc;
    uint64_t bits = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (wires[i] >= net_.n_wires()) return B2_ERR_INVALID_WIRE_ID;
        bits += net_.width[wires[i]];
    }
    if (bits > 0xFFFFFFF0ull) return B2_ERR_RANGE;
    if ((rc = materialise_valid())) return rc;
    const size_t n_bytes = (size_t)(bits / 8 + 1);  // always at least one unused bit (netcode lib.rs:168)
    if (cap < n_bytes) return B2_ERR_RANGE;
    if (bit_len) *bit_len = bits;
    std::vector<uint32_t> rows;
    rows.reserve((size_t)bits);
    for (uint32_t i = 0; i < n; i++)
        for (uint32_t j = 0; j < net_.width[wires[i]]; j++) rows.push_back(net_.row_of_bit[net_.bit_off[wires[i]] + j]);
    if ((rc = ensure_rows_buffer((uint32_t)std::max<size_t>(rows.size(), 1)))) return rc;
    rows_key_ = ptrs_key_ = 0;
    if (pack_cap_ < 2 * n_bytes) {
        CU_TRY(cudaStreamSynchronize(ST));
        cudaFree(d_pack_);
        d_pack_ = nullptr;
        pack_cap_ = 0;
        CU_TRY(cudaMalloc((void **)&d_pack_, 2 * n_bytes + 64));
        pack_cap_ = 2 * n_bytes + 64;
    }
    if (!rows.empty()) CU_TRY(cudaMemcpyAsync(d_rows_, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, ST));
    k_pack_lane0<<<(uint32_t)((n_bytes + 255) / 256), 256, 0, ST>>>(d_rows_, (uint32_t)bits, val_[cur_], vld_[cur_], stride_, d_pack_,
                                                                    d_pack_ + n_bytes, (uint32_t)n_bytes);
    LAUNCHED();
    CU_TRY(cudaMemcpyAsync(p0, d_pack_, n_bytes, cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaMemcpyAsync(p1, d_pack_ + n_bytes, n_bytes, cudaMemcpyDeviceToHost, ST));
    CU_TRY(cudaStreamSynchronize(ST));
    return B2_OK;
}

}  // namespace b2
