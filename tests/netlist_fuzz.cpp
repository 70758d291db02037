// Host-only fuzz of the netlist compiler (digilogic_b200/csrc/netlist.cpp), built with -fsanitize=address,undefined by
// tests/test_capi.py: random builder calls over every component kind (valid and invalid), then compile() and a check
// of the invariants the kernels rely on — every row index in range, implicit output rows, level ranges, resolve tables.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../digilogic_b200/csrc/netlist.hpp"
#include "../include/digilogic_b200.h"

using namespace b2;

#define REQUIRE(c)                                                     \
    do {                                                               \
        if (!(c)) {                                                    \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? std::atoi(argv[1]) : 200;
    {   // capacity: bit totals are counted in 64 bits and refused at the limits, never wrapped
        Builder b;
        uint32_t first = 0, cell = 0;
        const uint32_t n_fit = (uint32_t)(Builder::kMaxBits / 255);
        REQUIRE(b.add_wires(255, n_fit, &first) == B2_OK && first == 0);
        REQUIRE(b.total_bits == (uint64_t)n_fit * 255);
        REQUIRE(b.add_wires(255, 1, &first) == B2_ERR_OUT_OF_WIRES);       // 2^30 bits reached
        REQUIRE(b.add_wires(255, 0xFFFFFFFFu, &first) == B2_ERR_OUT_OF_WIRES);
        REQUIRE(b.width.size() == n_fit);
        uint32_t added = 0;
        int rc = B2_OK;
        while ((rc = b.add_arith(K_ADD, 0, 1, 2, &cell)) == B2_OK) added++;  // 255 x 510 bit-blasted inputs each
        REQUIRE(rc == B2_ERR_TOO_MANY_COMPONENTS);
        REQUIRE(added == Builder::kMaxBitInputs / (255ull * 510ull));
        REQUIRE(b.total_bit_inputs <= Builder::kMaxBitInputs && b.total_bit_gates <= Builder::kMaxBitGates);
        const uint32_t in8[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        while ((rc = b.add_gate(K_AND, in8, 8, 9, &cell)) == B2_OK) added++;  // 255 x 8 more each, up to the same limit
        REQUIRE(rc == B2_ERR_TOO_MANY_COMPONENTS && b.comps.size() == added);
        REQUIRE(b.total_bit_inputs <= Builder::kMaxBitInputs && b.total_bit_inputs + 255 * 8 > Builder::kMaxBitInputs);
    }
    std::mt19937 rng(12345);
    auto rnd = [&](uint32_t n) { return (uint32_t)(rng() % n); };
    for (int r = 0; r < rounds; r++) {
        Builder b;
        const uint32_t n_wires = 4 + rnd(40);
        for (uint32_t i = 0; i < n_wires; i++) {
            uint32_t id;
            const uint8_t w = rnd(4) == 0 ? (uint8_t)(1 + rnd(40)) : (uint8_t)(1 + rnd(3));
            REQUIRE(b.add_wire(w, &id) == B2_OK && id == i);
        }
        const uint32_t n_calls = rnd(120);
        for (uint32_t k = 0; k < n_calls; k++) {
            uint32_t in[8], cell;
            const uint32_t n = rnd(6);
            for (uint32_t i = 0; i < 8; i++) in[i] = rnd(n_wires + 1);  // sometimes one past the end
            const uint32_t out = rnd(n_wires + 1), x = rnd(n_wires + 1), y = rnd(n_wires + 1), z = rnd(n_wires + 1);
            switch (rnd(13)) {
                case 10: b.add_arith(K_MUL, x, y, out, &cell); break;
                case 11: b.add_compare((int)rnd(11), x, y, out, &cell); break;
                case 12: b.add_shift(K_SHL + (int)rnd(3), x, y, out, &cell); break;
                case 0: b.add_not(x, out, &cell); break;
                case 1: b.add_mux(in, n, x, out, &cell); break;
                case 2: b.add_buffer(x, y, out, &cell); break;
                case 3: b.add_slice(x, rnd(8), out, &cell); break;
                case 4: b.add_merge(in, n, out, &cell); break;
                case 5: b.add_register(x, out, y, z, &cell); break;
                case 6: b.add_arith(rnd(2) ? K_ADD : K_SUB, x, y, out, &cell); break;
                case 7: b.add_horizontal_gate((int)rnd(6), x, out, &cell); break;
                default: b.add_gate((int)rnd(6), in, n, out, &cell); break;
            }
        }
        Compiled c;
        compile(b, c);
        REQUIRE(c.n_rows >= c.n_src);
        REQUIRE(c.fan0.size() == c.kind2.size() && c.fan1.size() == c.kind2.size() && c.fan2.size() == c.kind2.size());
        REQUIRE(c.n_two <= c.kind2.size());
        REQUIRE(c.n_rows == c.n_src + c.n_fast() + c.n_wide() + c.n_res());
        for (uint32_t g = 0; g < c.n_fast(); g++) REQUIRE(c.fan0[g] < c.n_rows && c.fan1[g] < c.n_rows && c.fan2[g] < c.n_rows);
        REQUIRE(c.wide_off.size() == (size_t)c.n_wide() + 1 && c.wide_off.back() == c.wide_in.size());
        REQUIRE(c.wide_aux.size() == c.n_wide());
        for (uint32_t v : c.wide_in) REQUIRE(v < c.n_rows);
        for (uint32_t g = 0; g < c.n_wide(); g++) {
            const uint32_t n_in = c.wide_off[g + 1] - c.wide_off[g];
            if (c.wide_kind[g] == K_REG) REQUIRE(n_in == 4);
            if (c.wide_kind[g] == K_ADD || c.wide_kind[g] == K_SUB || c.wide_kind[g] == K_MUL) REQUIRE(n_in % 2 == 0 && c.wide_nsel[g] < n_in / 2);
            if (c.wide_kind[g] == K_CMP) REQUIRE(n_in % 2 == 0 && n_in >= 2 && c.wide_aux[g] <= CMP_GES);
            if (c.wide_kind[g] >= K_SHL && c.wide_kind[g] <= K_ASHR) REQUIRE(c.wide_aux[g] <= 8 && n_in > c.wide_aux[g] && c.wide_nsel[g] < n_in - c.wide_aux[g]);
            if (c.wide_kind[g] == K_MUX) REQUIRE(n_in == c.wide_nsel[g] + (1u << c.wide_nsel[g]));
        }
        REQUIRE(c.row_of_bit.size() == c.bit_off.back());
        for (uint32_t v : c.row_of_bit) REQUIRE(v < c.n_rows);
        if (c.n_res()) {
            REQUIRE(c.has_tristate && c.cyclic && c.res_off.back() == c.res_in.size());
            for (uint32_t v : c.res_in) REQUIRE(v >= c.n_src && v < c.n_src + c.n_fast() + c.n_wide());
        }
        for (size_t i = 0; i < c.res_of_bit.size(); i++)
            if (c.res_of_bit[i] >= 0) REQUIRE(c.row_of_bit[i] == c.n_src + c.n_fast() + c.n_wide() + (uint32_t)c.res_of_bit[i]);
        if (c.has_state || c.has_tristate) REQUIRE(c.cyclic);
        if (!c.cyclic) {
            REQUIRE(c.levels.size() == c.depth);
            uint32_t k2 = 0, k3 = c.n_two, kw = 0;
            for (const Level &lv : c.levels) {
                REQUIRE(lv.g2_begin == k2 && lv.g2_end >= k2 && lv.g3_begin == k3 && lv.g3_end >= k3 && lv.wide_begin == kw);
                REQUIRE(lv.g2_short >= lv.g2_begin && lv.g2_short <= lv.g2_end && lv.g3_short >= lv.g3_begin && lv.g3_short <= lv.g3_end);
                REQUIRE(lv.wide_short >= lv.wide_begin && lv.wide_short <= lv.wide_end);
                REQUIRE(lv.g2_never >= lv.g2_short && lv.g2_never <= lv.g2_end && lv.g3_never >= lv.g3_short && lv.g3_never <= lv.g3_end);
                REQUIRE(lv.wide_never >= lv.wide_short && lv.wide_never <= lv.wide_end);
                k2 = lv.g2_end; k3 = lv.g3_end; kw = lv.wide_end;
                // a gate only reads rows of earlier levels (or sources)
                for (uint32_t g = lv.g2_begin; g < lv.g2_end; g++) REQUIRE(c.fan0[g] < c.n_src + lv.g2_begin || c.fan0[g] >= c.n_src + c.n_two);
            }
            REQUIRE(k2 == c.n_two && k3 == c.n_fast() && kw == c.n_wide());
            // short-lived rows: nobody more than one level above reads them (the batch runner drops them from L2 after that level)
            std::vector<uint32_t> row_level(c.n_rows, 0), row_short(c.n_rows, 0), row_never(c.n_rows, 0);
            for (uint32_t l = 0; l < c.levels.size(); l++) {
                const Level &lv = c.levels[l];
                for (uint32_t g = lv.g2_begin; g < lv.g2_end; g++) { row_level[c.n_src + g] = l + 1; row_short[c.n_src + g] = g >= lv.g2_short; row_never[c.n_src + g] = g >= lv.g2_never; }
                for (uint32_t g = lv.g3_begin; g < lv.g3_end; g++) { row_level[c.n_src + g] = l + 1; row_short[c.n_src + g] = g >= lv.g3_short; row_never[c.n_src + g] = g >= lv.g3_never; }
                for (uint32_t g = lv.wide_begin; g < lv.wide_end; g++) {
                    row_level[c.n_src + c.n_fast() + g] = l + 1;
                    row_short[c.n_src + c.n_fast() + g] = g >= lv.wide_short;
                    row_never[c.n_src + c.n_fast() + g] = g >= lv.wide_never;
                }
            }
            auto reads_ok = [&](uint32_t reader_row, uint32_t r) { return !row_never[r] && (!row_short[r] || row_level[reader_row] <= row_level[r] + 1); };
            for (uint32_t g = 0; g < c.n_fast(); g++)
                REQUIRE(reads_ok(c.n_src + g, c.fan0[g]) && reads_ok(c.n_src + g, c.fan1[g]) && reads_ok(c.n_src + g, c.fan2[g]));
            for (uint32_t g = 0; g < c.n_wide(); g++)
                for (uint32_t i = c.wide_off[g]; i < c.wide_off[g + 1]; i++) REQUIRE(reads_ok(c.n_src + c.n_fast() + g, c.wide_in[i]));
        } else {
            REQUIRE(c.levels.size() <= 1);
        }
    }
    std::printf("netlist_fuzz ok (%d rounds)\n", rounds);
    return 0;
}
