// device_ops.cuh — device-side building blocks shared by the kernels of engine.cu and batch.cu: coherent row loads,
// the 4-state gate functions, the wide-gate evaluator (n-ary gates, multiplexers, registers, adder bits).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "netlist.hpp"

namespace b2 {

// ---------------------------------------------------------------------------- device helpers

__device__ __forceinline__ ulonglong2 ld_stream(const uint64_t *p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
// The same load; with `first` the line is inserted into L2 with evict-first priority (a one-off read of a row written long
// ago must not displace the rows the previous level has just written).  `first` is warp-uniform when a warp sits on one gate.
__device__ __forceinline__ ulonglong2 ld_stream_hint(const uint64_t *p, uint32_t first, uint64_t policy) {
    ulonglong2 r;
    if (first)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(policy));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
// Coherent variant for code that may run inside the device-side loop, where rows written by other CTAs
// earlier in the same launch are read back after a grid barrier: .nc data must be read-only for the launch, and
// a weak ld with L1::no_allocate was observed to return a stale row there (an SM re-reading a row it had read
// two applications earlier), so these loads are volatile — served from L2.
__device__ __forceinline__ ulonglong2 ld_row(const uint64_t *p) {
    ulonglong2 r;
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
// the same load, predicated off (result all-ones) when `skip`: no branch, no memory transaction
__device__ __forceinline__ ulonglong2 ld_stream_unless(const uint64_t *p, bool skip) {
    ulonglong2 r = make_ulonglong2(~0ull, ~0ull);
    asm volatile(
        "{ .reg .pred q; setp.eq.u32 q, %3, 0; @q ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2]; }"
        : "+l"(r.x), "+l"(r.y)
        : "l"(p), "r"((uint32_t)skip));
    return r;
}
// Row loads of a kernel launched with programmatic stream serialisation: the rows are written by the previous grid while this one
// is already resident, so they are NOT read-only for this kernel's lifetime and must not be .nc loads — ptxas is free to hoist a
// non-coherent load above griddepcontrol.wait (seen in SASS: LDG...CONSTANT before ACQBULK, and wrong results).  A plain weak
// load with a memory clobber stays behind the wait.
__device__ __forceinline__ ulonglong2 ld_dep(const uint64_t *p) {
    ulonglong2 r;
    asm volatile("ld.global.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
// with `first` the line is inserted into L2 with evict-first priority (see ld_stream_hint)
__device__ __forceinline__ ulonglong2 ld_dep_hint(const uint64_t *p, uint32_t first, uint64_t policy) {
    ulonglong2 r;
    if (first)
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(policy) : "memory");
    else
        asm volatile("ld.global.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ ulonglong2 ld_dep_unless(const uint64_t *p, bool skip) {
    ulonglong2 r = make_ulonglong2(~0ull, ~0ull);
    asm volatile(
        "{ .reg .pred q; setp.eq.u32 q, %3, 0; @q ld.global.L1::no_allocate.v2.u64 {%0,%1}, [%2]; }"
        : "+l"(r.x), "+l"(r.y)
        : "l"(p), "r"((uint32_t)skip)
        : "memory");
    return r;
}
__device__ __forceinline__ void st_pair(uint64_t *p, ulonglong2 v) {
    asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

// the same store; with `first` the line gets L2 evict-first priority (a row nobody reads soon)
__device__ __forceinline__ void st_pair_hint(uint64_t *p, ulonglong2 v, uint32_t first, uint64_t policy) {
    if (first) asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1,%2}, %3;" ::"l"(p), "l"(v.x), "l"(v.y), "l"(policy) : "memory");
    else asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

// 4-state gate on one 64-lane word: k1 = known one, k0 = known zero.
typedef unsigned long long u64;  // the element type of ulonglong2
// sem: SemFlags of the netlist — the reconstructed points A.6-6 / A.6-7 (does an enabled buffer / a slice or merge pass a
// High-Z input bit on as High-Z, or as X).
__device__ __forceinline__ void gate_word(uint32_t kind, u64 av, u64 am, u64 bv, u64 bm, u64 &rv, u64 &rm, uint32_t sem = 0) {
    u64 v, m;
    switch (kind) {
        case K_AND:
        case K_NAND: {
            const u64 k1 = (av & am) & (bv & bm), k0 = (~av & am) | (~bv & bm);
            m = k0 | k1;
            v = k1;
        } break;
        case K_OR:
        case K_NOR: {
            const u64 k1 = (av & am) | (bv & bm), k0 = (~av & am) & (~bv & bm);
            m = k0 | k1;
            v = k1;
        } break;
        case K_XOR:
        case K_XNOR:
            m = am & bm;
            v = av ^ bv;
            break;
        case K_BUF: {  // a = data, b = enable.  enable 1: the data (High-Z reads X); 0: High-Z; Z/X: X
            const u64 en1 = bv & bm;
            rm = en1 & am;
            rv = (en1 & ((sem & SEM_BUF_Z) ? av : (av | ~am))) | ~bm;
            return;
        }
        case K_RAW:  // hidden raw copy (a register's clock one step ago): High-Z stays High-Z
            rm = am;
            rv = av;
            return;
        case K_COPY:  // one bit of a slice / merge: the input, High-Z read as X (A.6-7 alternative: passed on as High-Z)
            rm = am;
            rv = (sem & SEM_COPY_Z) ? av : (av | ~am);
            return;
        default:  // K_NOT (single input)
            m = am;
            v = av;
            break;
    }
    if (kind >= K_NAND) v = ~v;
    rv = v | ~m;  // not-valid lanes carry the canonical X = (1,0)
    rm = m;
}

// One register bit (gsim add_register restated): rising edge = the clock was a valid 0 one step ago
// (pc) and is a valid 1 now (c); enable 1 loads the data bit (High-Z reads X), enable 0 keeps, enable Z/X makes the stored
// value Undefined.  The stored value is the register's own row (o); a row that is still High-Z has never been written and
// stands for the initial Undefined.
__device__ __forceinline__ void reg_word(u64 dv, u64 dm, u64 ev, u64 em, u64 cv, u64 cm, u64 pcv, u64 pcm, u64 ov, u64 om, u64 &rv,
                                         u64 &rm) {
    const u64 edge = pcm & ~pcv & cm & cv;
    const u64 take = edge & ev & em, keep = ~edge | (edge & ~ev & em), undef = edge & ~em;
    rm = (take & dm) | (keep & om);
    rv = (take & (dv | ~dm)) | (keep & (ov | ~om)) | undef;
}

// Multiplier, comparator and shifter bits (gsim add_mul, add_compare_*, add_*_shift restated) on one
// 64-lane word.  val(i) / vld(i): value / valid word of the gate's i-th input entry (layout: see netlist.hpp Kind).  Any input
// bit that is not a valid 0/1 makes the result Undefined, like add / sub.  Correct for every width up to 255; the multiplier
// is cubic in the width per wire (a bit-sliced column adder per result bit) — these components are not on any hot path.
template <class FV, class FM>
__device__ __forceinline__ void arith_word(uint32_t kind, uint32_t n_in, uint32_t nsel, uint32_t aux, FV val, FM vld, u64 &rv, u64 &rm) {
    u64 ok = ~0ull;
    for (uint32_t i = 0; i < n_in; i++) ok &= vld(i);
    u64 r = 0;
    if (kind == K_MUL) {
        const uint32_t w = n_in / 2, j = nsel;
        u64 cnt[10];  // bit-sliced counter of the current column: sum of its partial products + the carries of the columns below
#pragma unroll
        for (int k = 0; k < 10; k++) cnt[k] = 0;
        for (uint32_t c = 0; c <= j; c++) {
            for (uint32_t i = 0; i <= c; i++) {
                u64 carry = val(i) & val(w + c - i);
#pragma unroll
                for (int k = 0; k < 10; k++) {
                    const u64 t = cnt[k] & carry;
                    cnt[k] ^= carry;
                    carry = t;
                }
            }
            if (c == j) r = cnt[0];
#pragma unroll
            for (int k = 0; k < 9; k++) cnt[k] = cnt[k + 1];
            cnt[9] = 0;
        }
    } else if (kind == K_CMP) {
        const uint32_t w = n_in / 2;
        u64 lt = 0, eq = ~0ull;  // unsigned, from the top bit down
        for (uint32_t t = w; t-- > 0;) {
            const u64 x = val(t), y = val(w + t);
            lt |= eq & ~x & y;
            eq &= ~(x ^ y);
        }
        if (aux >= CMP_LTS) {  // signs differ: the negative operand is the smaller one
            const u64 sx = val(w - 1), diff = sx ^ val(2 * w - 1);
            lt = (diff & sx) | (~diff & lt);
        }
        switch (aux) {
            case CMP_EQ: r = eq; break;
            case CMP_NE: r = ~eq; break;
            case CMP_LTU:
            case CMP_LTS: r = lt; break;
            case CMP_GTU:
            case CMP_GTS: r = ~lt & ~eq; break;
            case CMP_LEU:
            case CMP_LES: r = lt | eq; break;
            default: r = ~lt; break;
        }
    } else {  // K_SHL / K_LSHR / K_ASHR: result bit j = the source bit of whichever amount the lane holds
        const uint32_t k = aux, w = n_in - k, j = nsel;
        const u64 sign = kind == K_ASHR ? val(w - 1) : 0;
        for (uint32_t amt = 0; amt < (1u << k); amt++) {
            u64 match = ~0ull;
            for (uint32_t s = 0; s < k; s++) match &= ((amt >> s) & 1) ? val(w + s) : ~val(w + s);
            u64 src;
            if (kind == K_SHL) src = j >= amt ? val(j - amt) : 0;
            else src = j + amt < w ? val(j + amt) : sign;
            r |= match & src;
        }
    }
    rm = ok;
    rv = r | ~ok;
}

// Driver conflict between two operands of a wire: both not High-Z; with `lenient` (the A.6-2
// alternative) two valid operands that agree do not conflict.
__device__ __forceinline__ u64 conflict_of(u64 av, u64 am, u64 bv, u64 bm, int lenient) {
    u64 c = (av | am) & (bv | bm);
    if (lenient) c &= ~(am & bm & ~(av ^ bv));
    return c;
}

__device__ __forceinline__ uint64_t lane_mask_of(uint32_t w, uint64_t n_lanes) {
    const uint64_t lo = (uint64_t)w * 64;
    if (lo + 64 <= n_lanes) return ~0ull;
    if (lo >= n_lanes) return 0ull;
    return (1ull << (n_lanes - lo)) - 1;
}

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}


struct WideArgs {
    const uint32_t *off, *in;
    const uint8_t *kind, *nsel;
    const uint64_t *in_val, *in_vld;
    uint64_t *out_val, *out_vld;
    uint64_t *changed;
    uint32_t g_begin, g_end, row_base, stride, pairs, pair_blocks;
    uint8_t *kv;  // nullable: rows whose valid plane is elided (read as all-ones); outputs are always stored
    const uint8_t *chg_prev;  // unit-delay frontier, as in Eval2Args
    uint8_t *chg_cur;
    const uint32_t *all_valid;  // as in Eval2Args (nullable)
    uint32_t sem = 0;           // SemFlags of the netlist
    uint32_t *p2diff = nullptr; // device-side loop: period-2 detector (see Eval2Args)
    const uint8_t *aux = nullptr; // per gate: comparator op / width of a shift amount
};

__device__ __forceinline__ ulonglong2 ld_valid(const WideArgs &a, uint32_t row, size_t col) {
    if (a.kv && (*a.all_valid || a.kv[row])) return make_ulonglong2(~0ull, ~0ull);
    return ld_row(a.in_vld + (size_t)row * a.stride + col);
}

// n-ary gates (left fold, inverted kinds invert once at the end) and multiplexer slices.
template <bool JACOBI>
__device__ __forceinline__ void eval_wide_gate(const WideArgs &a, uint32_t pb, uint32_t g, uint32_t x = 0xFFFFFFFFu) {
    const uint32_t pair = x != 0xFFFFFFFFu ? x : pb * blockDim.x + threadIdx.x;
    if (pair >= a.pairs || g >= a.g_end) return;
    const size_t col = (size_t)pair * 2;
    const uint32_t b = a.off[g], e = a.off[g + 1];
    const uint32_t kind = a.kind[g];
    if (JACOBI && a.chg_prev) {
        uint32_t act = a.chg_prev[a.row_base + g];
        for (uint32_t i = b; i < e; i++) act |= a.chg_prev[a.in[i]];
        if (!act) return;  // warp-uniform when a warp sits on one gate
    }
    ulonglong2 rv, rm;
    if (kind == K_ADD || kind == K_SUB) {
        // one bit of a + b or a - b: inputs a_0..a_{w-1}, b_0..b_{w-1}; any invalid input bit makes the result Undefined
        const uint32_t w = (e - b) / 2, j = a.nsel[g];
        const bool sub = kind == K_SUB;
        ulonglong2 ok = make_ulonglong2(~0ull, ~0ull), carry = sub ? ok : make_ulonglong2(0, 0), sum = make_ulonglong2(0, 0);
        for (uint32_t t = 0; t < w; t++) {
            const ulonglong2 ma = ld_valid(a, a.in[b + t], col), mb = ld_valid(a, a.in[b + w + t], col);
            ok.x &= ma.x & mb.x;
            ok.y &= ma.y & mb.y;
            if (t <= j) {
                const ulonglong2 x = ld_row(a.in_val + (size_t)a.in[b + t] * a.stride + col);
                ulonglong2 y = ld_row(a.in_val + (size_t)a.in[b + w + t] * a.stride + col);
                if (sub) {
                    y.x = ~y.x;
                    y.y = ~y.y;
                }
                if (t == j) {
                    sum.x = x.x ^ y.x ^ carry.x;
                    sum.y = x.y ^ y.y ^ carry.y;
                } else {
                    carry.x = (x.x & y.x) | (carry.x & (x.x ^ y.x));
                    carry.y = (x.y & y.y) | (carry.y & (x.y ^ y.y));
                }
            }
        }
        rm = ok;
        rv.x = sum.x | ~ok.x;
        rv.y = sum.y | ~ok.y;
    } else if (kind >= K_MUL && kind <= K_ASHR) {
        const uint32_t n_in = e - b, nsel = a.nsel[g], aux = a.aux ? a.aux[g] : 0;
        auto vx = [&](uint32_t i) { return (u64) * (const volatile uint64_t *)(a.in_val + (size_t)a.in[b + i] * a.stride + col); };
        auto vy = [&](uint32_t i) { return (u64) * (const volatile uint64_t *)(a.in_val + (size_t)a.in[b + i] * a.stride + col + 1); };
        auto mx = [&](uint32_t i) { return (u64)ld_valid(a, a.in[b + i], col).x; };
        auto my = [&](uint32_t i) { return (u64)ld_valid(a, a.in[b + i], col).y; };
        arith_word(kind, n_in, nsel, aux, vx, mx, rv.x, rm.x);
        arith_word(kind, n_in, nsel, aux, vy, my, rv.y, rm.y);
    } else if (kind == K_REG) {
        const size_t rd = (size_t)a.in[b] * a.stride + col, re = (size_t)a.in[b + 1] * a.stride + col,
                     rc = (size_t)a.in[b + 2] * a.stride + col, rp = (size_t)a.in[b + 3] * a.stride + col,
                     ro = (size_t)(a.row_base + g) * a.stride + col;
        const ulonglong2 dv = ld_row(a.in_val + rd), dm = ld_row(a.in_vld + rd), ev = ld_row(a.in_val + re), em = ld_row(a.in_vld + re),
                         cv = ld_row(a.in_val + rc), cm = ld_row(a.in_vld + rc), pv = ld_row(a.in_val + rp), pm = ld_row(a.in_vld + rp),
                         ov = ld_row(a.in_val + ro), om = ld_row(a.in_vld + ro);
        reg_word(dv.x, dm.x, ev.x, em.x, cv.x, cm.x, pv.x, pm.x, ov.x, om.x, rv.x, rm.x);
        reg_word(dv.y, dm.y, ev.y, em.y, cv.y, cm.y, pv.y, pm.y, ov.y, om.y, rv.y, rm.y);
    } else if (kind == K_MUX) {
        const uint32_t ns = a.nsel[g];
        ulonglong2 allv = make_ulonglong2(~0ull, ~0ull);
        for (uint32_t s = 0; s < ns; s++) {
            const ulonglong2 m = ld_valid(a, a.in[b + s], col);
            allv.x &= m.x;
            allv.y &= m.y;
        }
        ulonglong2 pv = make_ulonglong2(0, 0), pm = make_ulonglong2(0, 0);
        for (uint32_t j = 0; b + ns + j < e; j++) {
            ulonglong2 match = make_ulonglong2(~0ull, ~0ull);
            for (uint32_t s = 0; s < ns; s++) {
                const ulonglong2 sv = ld_row(a.in_val + (size_t)a.in[b + s] * a.stride + col);
                const uint64_t want = ((j >> s) & 1) ? ~0ull : 0ull;
                match.x &= ~(sv.x ^ want);
                match.y &= ~(sv.y ^ want);
            }
            const size_t r = (size_t)a.in[b + ns + j] * a.stride + col;
            const ulonglong2 dv = ld_row(a.in_val + r), dm = ld_valid(a, a.in[b + ns + j], col);
            pv.x |= match.x & dv.x;
            pv.y |= match.y & dv.y;
            pm.x |= match.x & dm.x;
            pm.y |= match.y & dm.y;
        }
        rm.x = allv.x & pm.x;
        rm.y = allv.y & pm.y;
        // High-Z on the selected input reads X (A.6-4 alternative: passes through as High-Z); invalid select -> X
        const bool zpass = (a.sem & SEM_MUX_Z) != 0;
        rv.x = (allv.x & pv.x) | (zpass ? ~allv.x : ~rm.x);
        rv.y = (allv.y & pv.y) | (zpass ? ~allv.y : ~rm.y);
    } else {
        const uint32_t base = kind >= K_NAND ? kind - 3 : kind;  // fold with the positive op
        size_t r = (size_t)a.in[b] * a.stride + col;
        rv = ld_row(a.in_val + r);
        rm = ld_valid(a, a.in[b], col);
        for (uint32_t i = b + 1; i < e; i++) {
            r = (size_t)a.in[i] * a.stride + col;
            const ulonglong2 xv = ld_row(a.in_val + r), xm = ld_valid(a, a.in[i], col);
            gate_word(base, rv.x, rm.x, xv.x, xm.x, rv.x, rm.x, a.sem);
            gate_word(base, rv.y, rm.y, xv.y, xm.y, rv.y, rm.y, a.sem);
        }
        if (kind >= K_NAND) {
            rv.x = ~rv.x | ~rm.x;
            rv.y = ~rv.y | ~rm.y;
        }
    }
    const size_t o = (size_t)(a.row_base + g) * a.stride + col;
    if (JACOBI) {
        const ulonglong2 ov = ld_row(a.in_val + o), om = ld_row(a.in_vld + o);
        const uint64_t dx = (rv.x ^ ov.x) | (rm.x ^ om.x), dy = (rv.y ^ ov.y) | (rm.y ^ om.y);
        if (dx) atomicOr((unsigned long long *)&a.changed[col], (unsigned long long)dx);
        if (dy) atomicOr((unsigned long long *)&a.changed[col + 1], (unsigned long long)dy);
        if (a.chg_cur && (dx | dy)) a.chg_cur[a.row_base + g] = 1;
        if (a.p2diff) {
            const ulonglong2 pv = ld_row(a.out_val + o), pm = ld_row(a.out_vld + o);
            if ((rv.x ^ pv.x) | (rm.x ^ pm.x) | (rv.y ^ pv.y) | (rm.y ^ pm.y)) *a.p2diff = 1;
        }
    }
    st_pair(a.out_val + o, rv);
    if (a.kv && *a.all_valid) return;  // valid fan-ins give a valid output: its plane stays elided
    st_pair(a.out_vld + o, rm);
    if (a.kv && pb == 0 && threadIdx.x == 0) a.kv[a.row_base + g] = 0;
}


}  // namespace b2
