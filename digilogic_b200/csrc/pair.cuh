// pair.cuh — level-pair fusion of the levelised pass (B2_OPT_PAIR_FUSION).
//
// The per-level kernels move 48 B through L2 per gate per 64-lane word (two fan-in rows read, one row written), and in a
// deep netlist most of a level's rows are read by the level directly above.  Here two dependent gates are evaluated by the
// same thread: a launch holds ITEMS, an item is one gate whose fan-ins were all written by earlier launches (the parent)
// plus up to two gates that read the parent's output and, apart from it, only rows of earlier launches (its children).
// The parent's value reaches its children in registers — no load — and a parent row that nobody else reads and that is
// not a designated output is not stored at all.  Everything else is as in k_eval2: one thread per 128-bit pair of
// lane-words, all plane loads of an item issued before the first use, rows that die within a launch of being written dropped
// from L2 (discard.global.L2) once their last reader has run, consecutive launches chained by programmatic dependent launch
// (row loads behind griddepcontrol.wait: plain loads, device_ops.cuh ld_dep).  Measured on config 3: half the launches, 17 % fewer
// sectors between the SMs and L2, the same time as the per-level kernels.  Where it pays is a netlist with NARROW levels, whose
// per-level kernels are launch-bound: the 32x32 array multiplier of config 2 (32 gates per level on average, depth 184) settles
// 2^20 vectors in 0.55 ms instead of 0.76 ms.  B2_OPT_PAIR_FUSION = 2 (default) turns it on for such netlists only.
//
// plan_pairs() is the host side: a greedy as-soon-as-possible assignment of every gate to (launch, parent | child) in
// level order.  A gate that cannot be a child (two fan-ins written by the launch it would join, a parent that already has two
// children, a parent that is itself a child) becomes a parent of the next launch, so the plan is valid for any acyclic
// netlist of one/two-input gates; config 3 (depth 200) becomes 100 launches plus a few for the gates pushed back.
//
// The results are bit-identical to the per-level kernels on every row that is kept (tests/test_gpu_pair.py); rows that
// are not kept are undefined after a run, as with B2_BATCH_OPT_DISCARD alone.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "device_ops.cuh"
#include "netlist.hpp"
#include "pair_plan.hpp"

namespace b2 {

struct PairArgs {
    const uint4 *items;         // of this launch
    const uint32_t *disc_rows;  // of this launch
    uint32_t n_items, n_disc;
    uint64_t *val, *vld;
    uint32_t stride, pairs, pair_blocks, sem;
    uint32_t hints;    // rows whose next reader is launches away are stored with an L2 evict-first hint
    uint32_t discard;  // the store's rows are whole 128-byte lines (stride % 16 == 0): dead rows are dropped from L2
    uint32_t pdl;
};

__device__ __forceinline__ void discard_line(uint64_t *p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }

// blockDim = (BX, BY) as in k_eval2: threadIdx.x walks 128-bit pairs of lane-words, threadIdx.y items.
__global__ void __launch_bounds__(256, 5) k_eval_pair(const PairArgs a) {
    const uint32_t pb = blockIdx.x % a.pair_blocks, ib = blockIdx.x / a.pair_blocks;
    const uint32_t pair = pb * blockDim.x + threadIdx.x;
    if (a.pdl == 1) asm volatile("griddepcontrol.launch_dependents;");
    if (pair >= a.pairs) return;
    const size_t col = (size_t)pair * 2;
    const uint32_t i = min(ib * blockDim.y + threadIdx.y, a.n_items - 1);
    const bool live = ib * blockDim.y + threadIdx.y < a.n_items;
    // the item (and the row this thread drops in the prologue) does not depend on anything the previous launch writes:
    // loaded before the dependency wait
    const uint4 q0 = __ldg(a.items + 2 * (size_t)i), q1 = __ldg(a.items + 2 * (size_t)i + 1);
    const uint32_t lpr = a.stride >> 4;
    const uint32_t cta_threads = blockDim.x * blockDim.y;
    const uint32_t j0 = blockIdx.x * cta_threads + threadIdx.y * blockDim.x + threadIdx.x;
    // (by the CTAs that start first; launches drop far fewer than 296 x 256 lines)
    const bool drops = a.discard && blockIdx.x < 296u * (256u / cta_threads) && j0 < a.n_disc * lpr;
    const uint32_t drow = drops ? __ldg(a.disc_rows + j0 / lpr) : 0;
    if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t meta = q0.w;
    const bool c1 = meta & PM_C1, c2 = meta & PM_C2;
    const bool skip1 = !c1 || (meta & PM_C1_SELF), skip2 = !c2 || (meta & PM_C2_SELF);
    const uint64_t *v = a.val + col, *m = a.vld + col;
    // (ld_dep, not .nc loads: the rows were written by the previous grid while this one was already resident, and ptxas hoists
    // non-coherent loads above griddepcontrol.wait)
    const ulonglong2 av = ld_dep(v + (size_t)q0.x * a.stride), am = ld_dep(m + (size_t)q0.x * a.stride);
    const ulonglong2 bv = ld_dep(v + (size_t)q0.y * a.stride), bm = ld_dep(m + (size_t)q0.y * a.stride);
    const ulonglong2 xv = ld_dep_unless(v + (size_t)q1.x * a.stride, skip1), xm = ld_dep_unless(m + (size_t)q1.x * a.stride, skip1);
    const ulonglong2 yv = ld_dep_unless(v + (size_t)q1.z * a.stride, skip2), ym = ld_dep_unless(m + (size_t)q1.z * a.stride, skip2);
    // a row whose last reader ran in the previous launch: dropped from L2 by the CTAs that start first, once this thread's
    // own loads are on their way
    if (drops) {
        const size_t w = (size_t)drow * a.stride + (j0 % lpr) * 16;
        discard_line(a.val + w);
        discard_line(a.vld + w);
    }
    uint64_t pol_first = 0;
    if (a.hints) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    const uint32_t late = a.hints ? meta : 0;
    ulonglong2 pv, pm;
    gate_word(meta & PM_KIND_MASK, av.x, am.x, bv.x, bm.x, pv.x, pm.x, a.sem);
    gate_word(meta & PM_KIND_MASK, av.y, am.y, bv.y, bm.y, pv.y, pm.y, a.sem);
    if (live) {
        const bool line_owner = a.discard && (pair & 7) == 0;
        if (meta & PM_STORE_P) {
            const size_t o = (size_t)q0.z * a.stride + col;
            st_pair_hint(a.val + o, pv, late & PM_LATE_P, pol_first);
            st_pair_hint(a.vld + o, pm, late & PM_LATE_P, pol_first);
            if ((meta & PM_NOW_P) && line_owner) {
                discard_line(a.val + o);
                discard_line(a.vld + o);
            }
        }
        if (c1) {
            const ulonglong2 ov = (meta & PM_C1_SELF) ? pv : xv, om = (meta & PM_C1_SELF) ? pm : xm;
            const bool sw = meta & PM_C1_SWAP;
            const uint32_t kind = (meta >> 5) & PM_KIND_MASK;
            ulonglong2 rv, rm;
            gate_word(kind, sw ? ov.x : pv.x, sw ? om.x : pm.x, sw ? pv.x : ov.x, sw ? pm.x : om.x, rv.x, rm.x, a.sem);
            gate_word(kind, sw ? ov.y : pv.y, sw ? om.y : pm.y, sw ? pv.y : ov.y, sw ? pm.y : om.y, rv.y, rm.y, a.sem);
            const size_t o = (size_t)q1.y * a.stride + col;
            st_pair_hint(a.val + o, rv, late & PM_LATE_C1, pol_first);
            st_pair_hint(a.vld + o, rm, late & PM_LATE_C1, pol_first);
            if ((meta & PM_NOW_C1) && line_owner) {
                discard_line(a.val + o);
                discard_line(a.vld + o);
            }
        }
        if (c2) {
            const ulonglong2 ov = (meta & PM_C2_SELF) ? pv : yv, om = (meta & PM_C2_SELF) ? pm : ym;
            const bool sw = meta & PM_C2_SWAP;
            const uint32_t kind = (meta >> 10) & PM_KIND_MASK;
            ulonglong2 rv, rm;
            gate_word(kind, sw ? ov.x : pv.x, sw ? om.x : pm.x, sw ? pv.x : ov.x, sw ? pm.x : om.x, rv.x, rm.x, a.sem);
            gate_word(kind, sw ? ov.y : pv.y, sw ? om.y : pm.y, sw ? pv.y : ov.y, sw ? pm.y : om.y, rv.y, rm.y, a.sem);
            const size_t o = (size_t)q1.w * a.stride + col;
            st_pair_hint(a.val + o, rv, late & PM_LATE_C2, pol_first);
            st_pair_hint(a.vld + o, rm, late & PM_LATE_C2, pol_first);
            if ((meta & PM_NOW_C2) && line_owner) {
                discard_line(a.val + o);
                discard_line(a.vld + o);
            }
        }
    }
    if (a.pdl == 2) asm volatile("griddepcontrol.launch_dependents;");
}

}  // namespace b2
