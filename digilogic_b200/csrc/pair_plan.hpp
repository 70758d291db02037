// pair_plan.hpp — host side of the level-pair fusion (pair.cuh has the kernel and the description): a greedy
// as-soon-as-possible assignment of every gate to (launch, parent | child) in level order, the rows that need not be
// stored, and per launch the rows whose last reader ran in the launch before (dropped from L2 in its prologue).
// Pure host C++ (checked host-only by tests/netlist_fuzz.cpp).
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

#include "netlist.hpp"

namespace b2 {

// meta word of an item
enum PairMeta : uint32_t {
    PM_KIND_MASK = 31,  // parent kind bits 0-4, child 1 kind bits 5-9, child 2 kind bits 10-14
    PM_STORE_P = 1u << 15,   // the parent's row has a reader outside the item, or is kept
    PM_NOW_P = 1u << 16,     // nobody reads the parent's row: dropped from L2 right after the store
    PM_C1 = 1u << 17, PM_C2 = 1u << 18,        // child present
    PM_C1_SELF = 1u << 19, PM_C2_SELF = 1u << 20,  // the child's other operand is the parent too (one-input gates, both fan-ins the parent)
    PM_NOW_C1 = 1u << 21, PM_NOW_C2 = 1u << 22,
    PM_C1_SWAP = 1u << 23, PM_C2_SWAP = 1u << 24,  // the parent is the child's SECOND operand
    // the row's next reader is two or more launches away: it is on its way to HBM whatever happens, so it is stored with an L2
    // evict-first hint and leaves the cache to the rows that die young (those are dropped before they are ever written back)
    PM_LATE_P = 1u << 25, PM_LATE_C1 = 1u << 26, PM_LATE_C2 = 1u << 27
};

struct PairItem {  // 32 bytes: two 128-bit loads
    uint32_t pf0, pf1, prow, meta;
    uint32_t c1_other, c1_row, c2_other, c2_row;
};

struct PairLaunch {
    uint32_t item_begin, item_end;  // into the item array
    uint32_t disc_begin, disc_end;  // into the discard-row list: rows whose last reader ran in the previous launch
};

struct PairPlanHost {
    std::vector<PairItem> items;
    std::vector<PairLaunch> launches;
    std::vector<uint32_t> disc_rows;
    uint64_t n_children = 0, n_unstored = 0;
};

// keep_sorted: rows that must hold their value after the run (designated outputs).  Returns false if the netlist is not
// eligible (anything but one/two-input gates of an acyclic single-driver netlist).
// keep_all: every row is kept (a plain simulator run, whose wires can all be read back): nothing is dropped, every parent is stored;
// what remains is half the launches and the children's operand in registers.
inline bool plan_pairs(const Compiled &net, const std::vector<uint32_t> &keep_sorted, PairPlanHost &plan, bool keep_all = false) {
    const uint32_t G = net.n_g2(), n_src = net.n_src;
    if (net.cyclic || net.multi_driver || net.has_tristate || net.has_state || net.n_g3() || net.n_wide() || net.n_res() || G == 0)
        return false;
    std::vector<int32_t> launch(G, -1);
    std::vector<uint8_t> is_child(G, 0), n_child(G, 0);
    std::vector<uint32_t> parent(G, 0), child(2 * (size_t)G, 0);
    auto launch_of_row = [&](uint32_t row) -> int32_t { return row < n_src ? -1 : launch[row - n_src]; };
    int32_t n_launch = 0;
    for (uint32_t g = 0; g < G; g++) {  // gates are in level order: fan-ins come first
        const uint32_t a = net.fan0[g], b = net.fan1[g];
        const int32_t la = launch_of_row(a), lb = launch_of_row(b), k = std::max(la, lb);
        int32_t mine = k + 1;
        if (k >= 0) {
            // the fan-ins written by launch k: exactly one distinct row, produced by a parent with a free child slot
            const bool a_in = la == k, b_in = lb == k;
            if (!(a_in && b_in && a != b)) {
                const uint32_t p = (a_in ? a : b) - n_src;
                if (!is_child[p] && n_child[p] < 2) {
                    is_child[g] = 1;
                    parent[g] = p;
                    child[2 * (size_t)p + n_child[p]++] = g;
                    mine = k;
                }
            }
        }
        launch[g] = mine;
        n_launch = std::max(n_launch, mine + 1);
    }
    // readers of every gate row: the last launch that reads it, and whether anything but its own children does
    std::vector<int32_t> last_reader(G, -1), first_reader(G, 0x7FFFFFFF);  // first_reader: outside the item
    std::vector<uint8_t> outside(G, 0);
    for (uint32_t g = 0; g < G; g++) {
        const uint32_t f[2] = {net.fan0[g], net.fan1[g]};
        for (int i = 0; i < (f[0] == f[1] ? 1 : 2); i++) {
            if (f[i] < n_src) continue;
            const uint32_t p = f[i] - n_src;
            last_reader[p] = std::max(last_reader[p], launch[g]);
            if (!(is_child[g] && parent[g] == p)) {
                outside[p] = 1;
                first_reader[p] = std::min(first_reader[p], launch[g]);
            }
        }
    }
    auto kept = [&](uint32_t g) { return keep_all || std::binary_search(keep_sorted.begin(), keep_sorted.end(), n_src + g); };
    // items in launch order (counting sort of the parents; gate order inside a launch)
    std::vector<uint32_t> count(n_launch + 1, 0), dcount(n_launch + 1, 0);
    for (uint32_t g = 0; g < G; g++)
        if (!is_child[g]) count[launch[g] + 1]++;
    for (int32_t k = 0; k < n_launch; k++) count[k + 1] += count[k];
    plan.items.assign(count[n_launch], PairItem{});
    std::vector<uint32_t> fill(count.begin(), count.end() - 1);
    // rows to drop in the prologue of launch last_reader + 1
    std::vector<std::pair<int32_t, uint32_t>> drops;
    auto lifetime = [&](uint32_t g, bool stored, uint32_t now_flag, uint32_t late_flag, uint32_t &meta) {
        if (stored && last_reader[g] >= 0 && first_reader[g] >= launch[g] + 2) meta |= late_flag;
        if (!stored || kept(g)) return;
        // only rows that die young are worth dropping: they are still dirty in L2 (a row read for the last time many launches
        // after it was written went to HBM long ago, and a discard costs L2 bandwidth like the store it would cancel)
        if (last_reader[g] < 0) meta |= now_flag;
        else if (last_reader[g] <= launch[g] + 1 && last_reader[g] + 1 < n_launch) drops.emplace_back(last_reader[g] + 1, n_src + g);
    };
    for (uint32_t g = 0; g < G; g++) {
        if (is_child[g]) continue;
        PairItem it{};
        it.pf0 = net.fan0[g];
        it.pf1 = net.fan1[g];
        it.prow = n_src + g;
        uint32_t meta = net.kind2[g] & PM_KIND_MASK;
        const bool store_p = kept(g) || outside[g] || last_reader[g] < 0;  // a row nobody reads is still written (and dropped)
        if (store_p) meta |= PM_STORE_P;
        else plan.n_unstored++;
        lifetime(g, store_p, PM_NOW_P, PM_LATE_P, meta);
        for (uint32_t c = 0; c < n_child[g]; c++) {
            const uint32_t h = child[2 * (size_t)g + c];
            const uint32_t a = net.fan0[h], b = net.fan1[h];
            const bool self = a == b || (a == it.prow && b == it.prow);
            const bool swap = !self && b == it.prow;  // the parent is operand b
            const uint32_t other = self ? it.prow : (swap ? a : b);
            meta |= (uint32_t)(net.kind2[h] & PM_KIND_MASK) << (c ? 10 : 5);
            meta |= c ? PM_C2 : PM_C1;
            if (self) meta |= c ? PM_C2_SELF : PM_C1_SELF;
            if (swap) meta |= c ? PM_C2_SWAP : PM_C1_SWAP;
            (c ? it.c2_other : it.c1_other) = other;
            (c ? it.c2_row : it.c1_row) = n_src + h;
            lifetime(h, true, c ? PM_NOW_C2 : PM_NOW_C1, c ? PM_LATE_C2 : PM_LATE_C1, meta);
            plan.n_children++;
        }
        it.meta = meta;
        plan.items[fill[launch[g]]++] = it;
    }
    std::sort(drops.begin(), drops.end());
    plan.disc_rows.resize(drops.size());
    for (size_t i = 0; i < drops.size(); i++) {
        plan.disc_rows[i] = drops[i].second;
        dcount[drops[i].first + 1]++;
    }
    for (int32_t k = 0; k < n_launch; k++) dcount[k + 1] += dcount[k];
    plan.launches.resize(n_launch);
    for (int32_t k = 0; k < n_launch; k++) plan.launches[k] = PairLaunch{count[k], count[k + 1], dcount[k], dcount[k + 1]};
    return true;
}

}  // namespace b2
