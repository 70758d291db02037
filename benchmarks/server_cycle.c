/* The reference server's per-Eval cycle, through the SimServer mirror of the C ABI, in plain C:
 *   SetNetDrive for every input   (crates/digilogic_netcode/src/server.rs:447-451 -> set_net_drive)
 *   Eval{max_steps: 10_000}       (server.rs:452-456 -> eval; client.rs:223)
 *   get_net_state for EVERY net   (Adapter::sim_state, server.rs:375-383 — one call per net)
 * and, as the alternative a maintainer can switch to, the bulk b2_server_sim_state (server.rs:96 TODO).
 * Reads a netlist written by server_cycle.py:  "W G I" / I input ids / G lines "kind n in.. out".
 * Prints one JSON line.  Exit 77 = no CUDA device. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "digilogic_b200.h"

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "r");
    const int cycles = atoi(argv[2]);
    if (!f) return 2;
    unsigned W, G, I;
    if (fscanf(f, "%u %u %u", &W, &G, &I) != 3) return 2;
    uint32_t *inputs = malloc(sizeof(uint32_t) * (I ? I : 1)), *nets = malloc(sizeof(uint32_t) * W);
    for (unsigned i = 0; i < I; i++)
        if (fscanf(f, "%u", &inputs[i]) != 1) return 2;
    b2_server *sv = b2_server_new();
    const uint64_t cl = 1;
    b2_server_client_connected(sv, cl);
    if (b2_server_begin_build(sv, cl)) return 3;
    for (unsigned w = 0; w < W; w++)
        if (b2_server_add_net(sv, cl, 1, &nets[w])) return 3;
    for (unsigned g = 0; g < G; g++) {
        unsigned kind, n, out;
        uint32_t in[8];
        if (fscanf(f, "%u %u", &kind, &n) != 2 || n > 8) return 2;
        for (unsigned i = 0; i < n; i++)
            if (fscanf(f, "%u", &in[i]) != 1) return 2;
        if (fscanf(f, "%u", &out) != 1) return 2;
        uint32_t cell;
        int rc = kind == B2_NOT ? b2_server_add_not_gate(sv, cl, 1, in[0], out, &cell)
                                : b2_server_add_gate(sv, cl, (int)kind, 1, in, n, out, &cell);
        if (rc) return 3;
    }
    fclose(f);
    int rc = b2_server_end_build(sv, cl);
    if (rc) { fprintf(stderr, "end_build: %d (%s)\n", rc, b2_last_error()); return 3; }

    const uint8_t one = 1, zero = 0;
    uint64_t checksum = 0;
    double t_set = 0, t_eval = 0, t_get = 0, t_bulk = 0;
    uint8_t *p0 = malloc(W / 8 + 2), *p1 = malloc(W / 8 + 2);
    for (int c = -2; c < cycles; c++) { /* two warm-up cycles */
        double t0 = now();
        for (unsigned i = 0; i < I; i++) {
            const uint8_t v = ((c + 2 + i * 7) >> (i % 3)) & 1 ? one : zero;
            rc = b2_server_set_net_drive(sv, cl, inputs[i], &v, 1, &one, 1);
            if (rc) { fprintf(stderr, "set_net_drive: %d (%s)\n", rc, b2_last_error()); return rc == 1 ? 77 : 4; }
        }
        double t1 = now();
        rc = b2_server_eval(sv, cl, 10000);
        if (rc) { fprintf(stderr, "eval: %d (%s)\n", rc, b2_last_error()); return strstr(b2_last_error(), "no CUDA") ? 77 : 4; }
        double t2 = now();
        for (unsigned w = 0; w < W; w++) {
            uint8_t width;
            const uint8_t *a, *b;
            if (b2_server_get_net_state(sv, cl, nets[w], &width, &a, &b)) return 4;
            checksum = checksum * 3 + (a[0] & 1) + 2 * (b[0] & 1);
        }
        double t3 = now();
        uint64_t bit_len = 0;
        if (b2_server_sim_state(sv, cl, nets, W, p0, p1, W / 8 + 2, &bit_len)) return 4;
        double t4 = now();
        if (c >= 0) { t_set += t1 - t0; t_eval += t2 - t1; t_get += t3 - t2; t_bulk += t4 - t3; }
    }
    printf("{\"nets\": %u, \"gates\": %u, \"inputs\": %u, \"cycles\": %d, \"us_set_drives\": %.2f, \"us_eval\": %.2f, "
           "\"us_get_every_net\": %.2f, \"us_cycle\": %.2f, \"us_bulk_sim_state\": %.2f, \"checksum\": \"%llx\"}\n",
           W, G, I, cycles, 1e6 * t_set / cycles, 1e6 * t_eval / cycles, 1e6 * t_get / cycles,
           1e6 * (t_set + t_eval + t_get) / cycles, 1e6 * t_bulk / cycles, (unsigned long long)checksum);
    b2_server_free(sv);
    return 0;
}
