/* A plain-C client of the batch part of include/digilogic_b200.h: a 16-bit ripple-carry adder, 2-valued host stimulus
 * planes for 8192 lanes, a 2-slot batch (the level-sweep kernel), outputs back into host planes, every lane checked
 * against A + B computed here.  Then the same batch on simulator replicas (B2_BATCH_OPT_SWEEP = 0) and a generated-
 * stimulus run whose two paths must agree.  Exit 0 = ok, 77 = no CUDA device. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "digilogic_b200.h"

#define CHECK(cond)                                                                                          \
    do {                                                                                                     \
        if (!(cond)) {                                                                                       \
            fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, b2_last_error());          \
            return 1;                                                                                        \
        }                                                                                                    \
    } while (0)

#define NBITS 16
#define WORDS 128 /* 8192 lanes */

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

int main(void) {
    b2_builder *b = b2_builder_new();
    uint32_t a0, b0, s0, first, cell;
    CHECK(b2_add_wires(b, 1, NBITS, &a0) == B2_OK && b2_add_wires(b, 1, NBITS, &b0) == B2_OK);
    CHECK(b2_add_wires(b, 1, NBITS + 1, &s0) == B2_OK);
    uint32_t carry = 0;
    for (int i = 0; i < NBITS; i++) {
        uint32_t x, g, p;
        uint32_t in[2] = {a0 + i, b0 + i};
        CHECK(b2_add_wires(b, 1, 3, &first) == B2_OK);
        x = first; g = first + 1; p = first + 2;
        CHECK(b2_add_gate(b, B2_XOR, in, 2, x, &cell) == B2_OK);
        CHECK(b2_add_gate(b, B2_AND, in, 2, g, &cell) == B2_OK);
        if (i == 0) { /* half adder: sum = x, carry = g */
            uint32_t xx[2] = {x, x};
            CHECK(b2_add_gate(b, B2_OR, xx, 2, s0, &cell) == B2_OK);
            carry = g;
        } else {
            uint32_t sx[2] = {x, carry}, pc[2] = {x, carry}, nc;
            CHECK(b2_add_gate(b, B2_XOR, sx, 2, s0 + i, &cell) == B2_OK);
            CHECK(b2_add_gate(b, B2_AND, pc, 2, p, &cell) == B2_OK);
            CHECK(b2_add_wire(b, 1, &nc) == B2_OK);
            uint32_t oc[2] = {g, p};
            CHECK(b2_add_gate(b, B2_OR, oc, 2, nc, &cell) == B2_OK);
            carry = nc;
        }
    }
    uint32_t cc[2] = {carry, carry};
    CHECK(b2_add_gate(b, B2_OR, cc, 2, s0 + NBITS, &cell) == B2_OK);
    b2_sim *sim = b2_build(b);
    CHECK(sim != NULL);

    uint32_t inputs[2 * NBITS], outputs[NBITS + 1];
    for (int i = 0; i < NBITS; i++) { inputs[i] = a0 + i; inputs[NBITS + i] = b0 + i; }
    for (int i = 0; i <= NBITS; i++) outputs[i] = s0 + i;
    b2_batch *bt = NULL;
    CHECK(b2_batch_new(sim, inputs, 2 * NBITS, outputs, NBITS + 1, 32, 2, &bt) == B2_OK);
    CHECK(b2_batch_set_option(bt, B2_BATCH_OPT_SWEEP, 1) == B2_OK); /* first the level-sweep kernel */
    b2_sim_free(sim); /* the batch keeps the compiled netlist alive */
    b2_builder_free(b);

    static uint64_t in_v[2 * NBITS][WORDS], in_m[2 * NBITS][WORDS], out_v[NBITS + 1][WORDS], out_m[NBITS + 1][WORDS];
    static uint64_t out2_v[NBITS + 1][WORDS], out2_m[NBITS + 1][WORDS];
    for (int i = 0; i < 2 * NBITS; i++)
        for (int w = 0; w < WORDS; w++) { in_v[i][w] = rnd(); in_m[i][w] = ~0ull; }
    int rc = b2_batch_run_planes(bt, &in_v[0][0], &in_m[0][0], WORDS, 0, WORDS, 10000, &out_v[0][0], &out_m[0][0], WORDS, 0);
    if (rc == B2_ERR_NO_DEVICE) {
        fprintf(stderr, "no CUDA device: %s\n", b2_last_error());
        return 77;
    }
    CHECK(rc == B2_OK);
    int worst = -1;
    CHECK(b2_batch_wait(bt, &worst) == B2_OK && worst == B2_RUN_OK);
    b2_batch_info info;
    CHECK(b2_batch_get_info(bt, &info) == B2_OK && info.sweep == 1 && info.last_tiles == WORDS / 32 && info.n_slots == 2);
    for (int w = 0; w < WORDS; w++)
        for (int l = 0; l < 64; l++) {
            uint32_t A = 0, B = 0, S = 0;
            for (int i = 0; i < NBITS; i++) {
                A |= (uint32_t)((in_v[i][w] >> l) & 1) << i;
                B |= (uint32_t)((in_v[NBITS + i][w] >> l) & 1) << i;
            }
            for (int i = 0; i <= NBITS; i++) {
                S |= (uint32_t)((out_v[i][w] >> l) & 1) << i;
                CHECK(((out_m[i][w] >> l) & 1) == 1);
            }
            CHECK(S == A + B);
        }
    /* the same batch on simulator replicas */
    CHECK(b2_batch_set_option(bt, B2_BATCH_OPT_SWEEP, 0) == B2_OK);
    CHECK(b2_batch_run_planes(bt, &in_v[0][0], &in_m[0][0], WORDS, 0, WORDS, 10000, &out2_v[0][0], &out2_m[0][0], WORDS, 0) == B2_OK);
    CHECK(b2_batch_wait(bt, &worst) == B2_OK && worst == B2_RUN_OK);
    CHECK(memcmp(out_v, out2_v, sizeof(out_v)) == 0 && memcmp(out_m, out2_m, sizeof(out_m)) == 0);
    /* generated 4-state stimulus: both paths again */
    CHECK(b2_batch_run_generated(bt, 7, 1, 1000, WORDS, 10000, &out2_v[0][0], &out2_m[0][0], WORDS, 0) == B2_OK);
    CHECK(b2_batch_wait(bt, &worst) == B2_OK && worst == B2_RUN_OK);
    CHECK(b2_batch_set_option(bt, B2_BATCH_OPT_SWEEP, 1) == B2_OK);
    CHECK(b2_batch_run_generated(bt, 7, 1, 1000, WORDS, 10000, &out_v[0][0], &out_m[0][0], WORDS, 0) == B2_OK);
    CHECK(b2_batch_wait(bt, &worst) == B2_OK && worst == B2_RUN_OK);
    CHECK(memcmp(out_v, out2_v, sizeof(out_v)) == 0 && memcmp(out_m, out2_m, sizeof(out_m)) == 0);
    b2_batch_free(bt);
    printf("c_batch_smoke ok (%llu kernel launches)\n", (unsigned long long)b2_kernel_launches());
    return 0;
}
