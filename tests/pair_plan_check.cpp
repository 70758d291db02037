// Host-only check of the level-pair planner (digilogic_b200/csrc/pair_plan.hpp), built and run by tests/test_capi.py.
// Random acyclic netlists of one/two-input gates (NOT gates, gates with both inputs on one wire, wide fan-out, random kept
// rows) are compiled, planned, and the plan is EXECUTED on a scalar 4-state model that tracks which rows hold data:
//   * every gate is evaluated exactly once;
//   * a parent reads only rows written by earlier launches (or sources), a child's other operand likewise;
//   * no row is read after it was dropped, or without ever having been stored;
//   * kept rows end up equal to a plain gate-by-gate evaluation.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../digilogic_b200/csrc/netlist.hpp"
#include "../digilogic_b200/csrc/pair_plan.hpp"
#include "../include/digilogic_b200.h"

using namespace b2;

#define REQUIRE(c)                                                              \
    do {                                                                        \
        if (!(c)) {                                                             \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                           \
        }                                                                       \
    } while (0)

struct S4 {
    uint64_t v, m;
};

// SURVEY.md A.3 on one 64-lane word (the same table as device_ops.cuh gate_word, default switches)
static S4 gate(uint32_t kind, S4 a, S4 b) {
    uint64_t v, m;
    switch (kind) {
        case K_AND: case K_NAND: { const uint64_t k1 = (a.v & a.m) & (b.v & b.m), k0 = (~a.v & a.m) | (~b.v & b.m); m = k0 | k1; v = k1; } break;
        case K_OR: case K_NOR: { const uint64_t k1 = (a.v & a.m) | (b.v & b.m), k0 = (~a.v & a.m) & (~b.v & b.m); m = k0 | k1; v = k1; } break;
        case K_XOR: case K_XNOR: m = a.m & b.m; v = a.v ^ b.v; break;
        case K_COPY: return S4{a.v | ~a.m, a.m};
        default: m = a.m; v = a.v; break;  // K_NOT
    }
    if (kind >= K_NAND) v = ~v;
    return S4{v | ~m, m};
}

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? std::atoi(argv[1]) : 300;
    std::mt19937_64 rng(0xD1610009);
    auto rnd = [&](uint32_t n) { return (uint32_t)(rng() % n); };
    uint64_t total_children = 0, total_unstored = 0, total_gates = 0;
    for (int round = 0; round < rounds; round++) {
        Builder b;
        const uint32_t n_in = 1 + rnd(12), n_gates = 1 + rnd(400);
        uint32_t w = 0, cell = 0;
        for (uint32_t i = 0; i < n_in; i++) b.add_wire(1, &w);
        const bool deep = rnd(2);  // deep: first input from the recent past (long chains), else anywhere
        for (uint32_t g = 0; g < n_gates; g++) {
            const uint32_t have = n_in + g;
            b.add_wire(1, &w);
            const uint32_t x = deep ? have - 1 - rnd(std::min(have, 3u)) : rnd(have);
            const uint32_t y = rnd(8) == 0 ? x : rnd(have);
            if (rnd(6) == 0) {
                REQUIRE(b.add_not(x, w, &cell) == B2_OK);
            } else {
                const uint32_t in[2] = {x, y};
                REQUIRE(b.add_gate((int)rnd(6), in, 2, w, &cell) == B2_OK);
            }
        }
        Compiled c;
        compile(b, c);
        REQUIRE(!c.cyclic && c.n_g3() == 0 && c.n_wide() == 0);
        const uint32_t G = c.n_g2(), n_src = c.n_src;
        std::vector<uint32_t> keep;
        for (uint32_t r = n_src; r < c.n_rows; r++)
            if (rnd(10) == 0) keep.push_back(r);
        for (int keep_all = 0; keep_all < 2; keep_all++) {
        PairPlanHost plan;
        REQUIRE(plan_pairs(c, keep, plan, keep_all != 0));
        if (keep_all) {
            REQUIRE(plan.n_unstored == 0 && plan.disc_rows.empty());
            for (const PairItem &it : plan.items) REQUIRE((it.meta & PM_STORE_P) && !(it.meta & (PM_NOW_P | PM_NOW_C1 | PM_NOW_C2)));
        }
        // plain evaluation
        std::vector<S4> ref(c.n_rows), st(c.n_rows, S4{0, 0});
        for (uint32_t r = 0; r < n_src; r++) ref[r] = st[r] = S4{rng(), rng() | rng()};
        for (uint32_t g = 0; g < G; g++) ref[n_src + g] = gate(c.kind2[g], ref[c.fan0[g]], ref[c.fan1[g]]);
        // the plan, executed
        std::vector<int32_t> written(c.n_rows, -2);  // launch that stored the row; -1 sources; -2 no data
        for (uint32_t r = 0; r < n_src; r++) written[r] = -1;
        std::vector<uint32_t> evaluated(G, 0);
        uint32_t items_seen = 0, disc_seen = 0;
        for (size_t k = 0; k < plan.launches.size(); k++) {
            const PairLaunch &L = plan.launches[k];
            REQUIRE(L.item_begin == items_seen && L.item_end > L.item_begin && L.disc_begin == disc_seen && L.disc_end >= L.disc_begin);
            items_seen = L.item_end;
            disc_seen = L.disc_end;
            for (uint32_t d = L.disc_begin; d < L.disc_end; d++) {
                const uint32_t r = plan.disc_rows[d];
                REQUIRE(r >= n_src && r < c.n_rows && written[r] >= 0 && written[r] < (int32_t)k);
                REQUIRE(!std::binary_search(keep.begin(), keep.end(), r));
                written[r] = -2;
            }
            auto readable = [&](uint32_t r) { return r < c.n_rows && written[r] >= -1 && written[r] < (int32_t)k; };
            for (uint32_t i = L.item_begin; i < L.item_end; i++) {
                const PairItem &it = plan.items[i];
                REQUIRE(readable(it.pf0) && readable(it.pf1));
                const uint32_t pg = it.prow - n_src;
                REQUIRE(it.prow >= n_src && pg < G && c.fan0[pg] == it.pf0 && c.fan1[pg] == it.pf1 && (it.meta & PM_KIND_MASK) == c.kind2[pg]);
                evaluated[pg]++;
                const S4 p = gate(it.meta & PM_KIND_MASK, st[it.pf0], st[it.pf1]);
                if (it.meta & PM_STORE_P) {
                    st[it.prow] = p;
                    written[it.prow] = (it.meta & PM_NOW_P) ? -2 : (int32_t)k;
                } else {
                    REQUIRE(!(it.meta & PM_NOW_P) && !std::binary_search(keep.begin(), keep.end(), it.prow));
                }
                for (int ch = 0; ch < 2; ch++) {
                    if (!(it.meta & (ch ? PM_C2 : PM_C1))) continue;
                    const uint32_t row = ch ? it.c2_row : it.c1_row, other = ch ? it.c2_other : it.c1_other;
                    const bool self = it.meta & (ch ? PM_C2_SELF : PM_C1_SELF), swap = it.meta & (ch ? PM_C2_SWAP : PM_C1_SWAP);
                    const uint32_t kind = (it.meta >> (ch ? 10 : 5)) & PM_KIND_MASK, hg = row - n_src;
                    REQUIRE(row >= n_src && hg < G && kind == c.kind2[hg]);
                    if (self) REQUIRE(c.fan0[hg] == it.prow && c.fan1[hg] == it.prow && !swap);
                    else REQUIRE(readable(other) && (swap ? (c.fan0[hg] == other && c.fan1[hg] == it.prow) : (c.fan0[hg] == it.prow && c.fan1[hg] == other)));
                    evaluated[hg]++;
                    const S4 o = self ? p : st[other];
                    st[row] = swap ? gate(kind, o, p) : gate(kind, p, o);
                    written[row] = (it.meta & (ch ? PM_NOW_C2 : PM_NOW_C1)) ? -2 : (int32_t)k;
                }
            }
        }
        REQUIRE(items_seen == plan.items.size() && disc_seen == plan.disc_rows.size());
        for (uint32_t g = 0; g < G; g++) REQUIRE(evaluated[g] == 1);
        for (uint32_t r : keep) REQUIRE(written[r] >= 0 && st[r].v == ref[r].v && st[r].m == ref[r].m);
        // every row that still holds data holds the right data
        for (uint32_t r = n_src; r < c.n_rows; r++)
            if (written[r] >= 0) REQUIRE(st[r].v == ref[r].v && st[r].m == ref[r].m);
        REQUIRE(plan.launches.size() <= c.depth);
        if (keep_all)
            for (uint32_t r = n_src; r < c.n_rows; r++) REQUIRE(written[r] >= 0);   // a plain run can read every wire back
        if (!keep_all) {
            total_children += plan.n_children;
            total_unstored += plan.n_unstored;
            total_gates += G;
        }
        }
    }
    REQUIRE(total_children > total_gates / 5);
    std::printf("pair_plan_check ok (%d rounds, %llu gates, %llu as children, %llu rows never stored)\n", rounds,
                (unsigned long long)total_gates, (unsigned long long)total_children, (unsigned long long)total_unstored);
    return 0;
}
