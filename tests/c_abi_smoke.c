/* A plain-C client of include/digilogic_b200.h: the calls GsimServer makes on gsim
 * (crates/digilogic_gsim/src/lib.rs:12-230), through b2_server_* and directly.
 * Exit code 0 = all checks passed; 77 = no CUDA device (the engine has no CPU path). */
#include <stdio.h>
#include <string.h>

#include "digilogic_b200.h"

#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, b2_last_error()); \
            return 1;                                                       \
        }                                                                   \
    } while (0)

int main(void) {
    /* builder validation is host-only */
    b2_builder *b = b2_builder_new();
    uint32_t a, c, o, cell;
    CHECK(b2_add_wire(b, 1, &a) == B2_OK && b2_add_wire(b, 1, &c) == B2_OK && b2_add_wire(b, 1, &o) == B2_OK);
    uint32_t ins[2] = {a, c};
    CHECK(b2_add_gate(b, B2_XOR, ins, 1, o, &cell) == B2_ERR_TOO_FEW_INPUTS);
    CHECK(b2_add_gate(b, B2_XOR, ins, 2, 9, &cell) == B2_ERR_INVALID_WIRE_ID);
    CHECK(b2_add_gate(b, B2_XOR, ins, 2, o, &cell) == B2_OK && cell == 0);
    b2_sim *s = b2_build(b);
    CHECK(s != NULL);
    b2_sim_info info;
    CHECK(b2_sim_get_info(s, &info) == B2_OK && info.n_gates2 == 1 && info.depth == 1 && !info.cyclic);
    uint32_t one = 1, zero = 0, p0 = 9, p1 = 9;
    int rc = b2_set_wire_drive(s, a, &one, &one, 1);
    if (rc == B2_ERR_NO_DEVICE) {
        fprintf(stderr, "no CUDA device: %s\n", b2_last_error());
        return 77;
    }
    CHECK(rc == B2_OK);
    CHECK(b2_set_wire_drive(s, c, &zero, &one, 1) == B2_OK);
    CHECK(b2_run_sim(s, 10000) == B2_RUN_OK);
    CHECK(b2_get_wire_state(s, o, &p0, &p1, 1) == B2_OK && p0 == 1 && p1 == 1); /* 1 xor 0 = 1 */
    CHECK(b2_set_wire_drive(s, c, &zero, &zero, 1) == B2_OK);                    /* High-Z input */
    CHECK(b2_run_sim(s, 10000) == B2_RUN_OK);
    CHECK(b2_get_wire_state(s, o, &p0, &p1, 1) == B2_OK && p0 == 1 && p1 == 0); /* X */
    b2_sim_free(s);
    b2_builder_free(b);

    /* two tri-state buffers on one bus (gsim add_buffer): hand-over, floating bus, conflict */
    b = b2_builder_new();
    uint32_t d0, d1, e0, e1, bus;
    CHECK(b2_add_wire(b, 1, &d0) == B2_OK && b2_add_wire(b, 1, &d1) == B2_OK && b2_add_wire(b, 1, &e0) == B2_OK);
    CHECK(b2_add_wire(b, 1, &e1) == B2_OK && b2_add_wire(b, 1, &bus) == B2_OK);
    CHECK(b2_add_buffer(b, d0, e0, 99, &cell) == B2_ERR_INVALID_WIRE_ID);
    CHECK(b2_add_buffer(b, d0, e0, bus, &cell) == B2_OK && b2_add_buffer(b, d1, e1, bus, &cell) == B2_OK);
    s = b2_build(b);
    CHECK(s != NULL && b2_sim_get_info(s, &info) == B2_OK && info.cyclic); /* tri-state nets: unit-delay path only */
    CHECK(b2_set_wire_drive(s, d0, &one, &one, 1) == B2_OK && b2_set_wire_drive(s, d1, &zero, &one, 1) == B2_OK);
    CHECK(b2_set_wire_drive(s, e0, &one, &one, 1) == B2_OK && b2_set_wire_drive(s, e1, &zero, &one, 1) == B2_OK);
    CHECK(b2_run_sim(s, 100) == B2_RUN_OK);
    CHECK(b2_get_wire_state(s, bus, &p0, &p1, 1) == B2_OK && p0 == 1 && p1 == 1);
    CHECK(b2_set_wire_drive(s, e0, &zero, &one, 1) == B2_OK && b2_set_wire_drive(s, e1, &one, &one, 1) == B2_OK);
    CHECK(b2_run_sim(s, 100) == B2_RUN_OK);
    CHECK(b2_get_wire_state(s, bus, &p0, &p1, 1) == B2_OK && p0 == 0 && p1 == 1);
    CHECK(b2_set_wire_drive(s, e1, &zero, &one, 1) == B2_OK);
    CHECK(b2_run_sim(s, 100) == B2_RUN_OK);
    CHECK(b2_get_wire_state(s, bus, &p0, &p1, 1) == B2_OK && p0 == 0 && p1 == 0); /* nobody drives: High-Z */
    CHECK(b2_set_wire_drive(s, e0, &one, &one, 1) == B2_OK && b2_set_wire_drive(s, e1, &one, &one, 1) == B2_OK);
    CHECK(b2_run_sim(s, 100) == B2_RUN_CONFLICT);
    b2_sim_free(s);
    b2_builder_free(b);

    /* the SimServer mirror: a NAND latch driven like the GUI client drives it */
    b2_server *sv = b2_server_new();
    const uint64_t id = 42;
    b2_server_client_connected(sv, id);
    CHECK(b2_server_begin_build(sv, id) == B2S_OK);
    uint32_t sn, rn, q, qn;
    CHECK(b2_server_add_net(sv, id, 1, &sn) == B2S_OK && b2_server_add_net(sv, id, 1, &rn) == B2S_OK);
    CHECK(b2_server_add_net(sv, id, 1, &q) == B2S_OK && b2_server_add_net(sv, id, 1, &qn) == B2S_OK);
    uint32_t g1[2] = {sn, qn}, g2[2] = {rn, q};
    CHECK(b2_server_add_gate(sv, id, B2_NAND, 1, g1, 2, q, &cell) == B2S_OK);
    CHECK(b2_server_add_gate(sv, id, B2_NAND, 2, g2, 2, qn, &cell) == B2S_WIDTH_MISMATCH);
    CHECK(b2_server_add_gate(sv, id, B2_NAND, 1, g2, 2, qn, &cell) == B2S_OK);
    CHECK(b2_server_eval(sv, id, 10) == B2S_INVALID_STATE);
    CHECK(b2_server_end_build(sv, id) == B2S_OK);
    const uint8_t hi = 1, lo = 0;
    CHECK(b2_server_set_net_drive(sv, id, sn, &lo, 1, &hi, 1) == B2S_OK); /* S' = 0 */
    CHECK(b2_server_set_net_drive(sv, id, rn, &hi, 1, &hi, 1) == B2S_OK); /* R' = 1 */
    CHECK(b2_server_eval(sv, id, 10000) == B2S_OK);
    uint8_t width;
    const uint8_t *s0, *s1;
    CHECK(b2_server_get_net_state(sv, id, q, &width, &s0, &s1) == B2S_OK && width == 1 && (s0[0] & 1) == 1 && (s1[0] & 1) == 1);
    CHECK(b2_server_set_net_drive(sv, id, sn, &lo, 1, &hi, 1) == B2S_OK);
    CHECK(b2_server_set_net_drive(sv, id, rn, &lo, 1, &hi, 1) == B2S_OK); /* S' = R' = 0 */
    CHECK(b2_server_eval(sv, id, 10000) == B2S_OK);
    CHECK(b2_server_set_net_drive(sv, id, sn, &hi, 1, &hi, 1) == B2S_OK);
    CHECK(b2_server_set_net_drive(sv, id, rn, &hi, 1, &hi, 1) == B2S_OK); /* both rise together: oscillation */
    CHECK(b2_server_eval(sv, id, 10000) == B2S_MAX_STEPS_REACHED);
    uint32_t nets[4] = {sn, rn, q, qn};
    uint8_t pl0[2], pl1[2];
    uint64_t bits = 0;
    CHECK(b2_server_sim_state(sv, id, nets, 4, pl0, pl1, sizeof(pl0), &bits) == B2S_OK && bits == 4);
    CHECK((pl1[0] & 0xF) == 0xF && (pl0[0] & 0x3) == 0x3);
    b2_server_free(sv);
    printf("c_abi_smoke ok (%s, %llu kernel launches)\n", b2_version(), (unsigned long long)b2_kernel_launches());
    return 0;
}
