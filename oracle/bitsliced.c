/*
 * bitsliced.c — second CPU model: 64 stimulus lanes per machine word, dense
 * synchronous unit-delay (Jacobi) evaluation of the same semantics as
 * gsim_oracle.c ("equivalent dense formulation").
 *
 * TEST INFRASTRUCTURE ONLY (see the header of gsim_oracle.c): used by tests/
 * to check large GPU runs lane-for-lane and by bench.py as the "best CPU"
 * context figure.  It is itself cross-checked lane-by-lane against the scalar
 * event-driven restatement in tests/test_oracle.py.  PARITY UNPINNED, as for
 * gsim_oracle.c.
 *
 * Model (1-bit wires only — the digilogic client only ever creates 1-bit
 * nets, reference crates/digilogic_netcode/src/client.rs:316):
 *   wire state W[w], external drive D[w]; component outputs are a pure
 *   function F(W) after the first begin_sim, so one application
 *       W' = G(W) :  W'[w] = resolve(D[w], { F_c(W) : c drives w })
 *   is one gsim step.  run_sim(max_steps) lets lane l perform applications
 *   k = 1..N+2 (N = max_steps+1; a cold start spends k = 1 on W_1 = D); the
 *   lane is Ok iff some k <= N+2 has W_k == W_{k-1}, else MaxStepsReached with
 *   state W_{N+1}; a driver conflict while computing W_k, k <= N+1, is Err.
 *   Nets with several drivers (tri-state buffers): gsim's two stop checks — "no wire changed" and "no
 *   component output changed" — no longer coincide (two drivers can swap roles and leave the wire as it
 *   was), so the component outputs O_k = F(W_{k-1}) are kept too and "changed" is W_k != W_{k-1} or
 *   O_k != O_{k-1}.  The chain  no output change at k  =>  no wire change at k  =>  no output change at
 *   k+1  makes "nothing changed at some k <= N+2" equal to gsim's last check (outputs, after step N).
 */
#include "gsim_oracle.h"
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t n_wires, n_comps, words;
    comp_t *comps;
    uint32_t *inputs;
    uint32_t *drivers_off, *drivers;
    uint64_t *dv, *dm;      /* drive planes  [wire][words] */
    uint64_t *av, *am;      /* state A       */
    uint64_t *bv, *bm;      /* state B (scratch) */
    uint64_t *ov, *om;      /* component outputs O = F(W) of the previous application [comp][words] */
    int multi;              /* some wire has two or more drivers */
    /* registers: stored value and previous clock (raw copy of the clock wire, so High-Z/X = "unknown"), current and
       next; the next ones are committed together with state B (never on the check-only application) */
    uint64_t *rdv, *rdm, *rpv, *rpm, *ndv, *ndm, *npv, *npm;
    int has_reg;
    int cold;
    int topo;               /* component order is topological and single-driver: DAG fast check allowed */
    uint64_t applications;  /* G applications performed by the last run */
} bs_sim;

static inline void f_and(uint64_t av, uint64_t am, uint64_t bv, uint64_t bm, uint64_t *rv, uint64_t *rm) {
    uint64_t k0a = ~av & am, k1a = av & am, k0b = ~bv & bm, k1b = bv & bm;
    uint64_t m = k0a | k0b | (k1a & k1b);
    *rm = m; *rv = (k1a & k1b) | ~m;
}
static inline void f_or(uint64_t av, uint64_t am, uint64_t bv, uint64_t bm, uint64_t *rv, uint64_t *rm) {
    uint64_t k0a = ~av & am, k1a = av & am, k0b = ~bv & bm, k1b = bv & bm;
    uint64_t m = k1a | k1b | (k0a & k0b);
    *rm = m; *rv = k1a | k1b | ~m;
}
static inline void f_xor(uint64_t av, uint64_t am, uint64_t bv, uint64_t bm, uint64_t *rv, uint64_t *rm) {
    uint64_t m = am & bm;
    *rm = m; *rv = (av ^ bv) | ~m;
}

/* Exposed for the truth-table test (A.3 k0/k1 formulas vs the gsim-style formulas). */
void bs_logic_op(int kind, uint64_t av, uint64_t am, uint64_t bv, uint64_t bm, uint64_t *rv, uint64_t *rm) {
    uint64_t v, m;
    switch (kind) {
    case GO_AND: case GO_NAND: f_and(av, am, bv, bm, &v, &m); break;
    case GO_OR: case GO_NOR: f_or(av, am, bv, bm, &v, &m); break;
    case GO_XOR: case GO_XNOR: f_xor(av, am, bv, bm, &v, &m); break;
    default: v = av; m = am; break;
    }
    if (kind >= GO_NAND && kind <= GO_NOT) v = ~v | ~m;
    *rv = v; *rm = m;
}

bs_sim *bs_new(const go_builder *b, uint32_t n_lanes) {
    for (uint32_t w = 0; w < b->n_wires; w++) if (b->width[w] != 1) return NULL;
    bs_sim *s = (bs_sim *)calloc(1, sizeof(bs_sim));
    uint32_t W = b->n_wires, C = b->n_comps;
    s->n_wires = W; s->n_comps = C; s->words = (n_lanes + 63) / 64;
    s->comps = (comp_t *)malloc((C ? C : 1) * sizeof(comp_t)); memcpy(s->comps, b->comps, (size_t)C * sizeof(comp_t));
    s->inputs = (uint32_t *)malloc((b->n_inputs ? b->n_inputs : 1) * sizeof(uint32_t));
    memcpy(s->inputs, b->inputs, (size_t)b->n_inputs * sizeof(uint32_t));
    s->drivers_off = (uint32_t *)calloc((size_t)W + 2, sizeof(uint32_t));
    for (uint32_t c = 0; c < C; c++) s->drivers_off[s->comps[c].output + 2]++;
    for (uint32_t w = 0; w < W; w++) s->drivers_off[w + 2] += s->drivers_off[w + 1];
    s->drivers = (uint32_t *)malloc((C ? C : 1) * sizeof(uint32_t));
    for (uint32_t c = 0; c < C; c++) s->drivers[s->drivers_off[s->comps[c].output + 1]++] = c;
    size_t n = (size_t)(W ? W : 1) * (s->words ? s->words : 1);
    s->dv = (uint64_t *)calloc(n, 8); s->dm = (uint64_t *)calloc(n, 8);
    s->av = (uint64_t *)calloc(n, 8); s->am = (uint64_t *)calloc(n, 8);
    s->bv = (uint64_t *)calloc(n, 8); s->bm = (uint64_t *)calloc(n, 8);
    size_t nc = (size_t)(C ? C : 1) * (s->words ? s->words : 1);
    s->ov = (uint64_t *)calloc(nc, 8); s->om = (uint64_t *)calloc(nc, 8);
    s->rdv = (uint64_t *)malloc(nc * 8); s->rdm = (uint64_t *)calloc(nc, 8); s->rpv = (uint64_t *)calloc(nc, 8); s->rpm = (uint64_t *)calloc(nc, 8);
    s->ndv = (uint64_t *)malloc(nc * 8); s->ndm = (uint64_t *)calloc(nc, 8); s->npv = (uint64_t *)calloc(nc, 8); s->npm = (uint64_t *)calloc(nc, 8);
    memset(s->rdv, 0xFF, nc * 8); memset(s->ndv, 0xFF, nc * 8);  /* stored value starts Undefined (1,0), clock history unknown */
    for (uint32_t c = 0; c < C; c++) if (s->comps[c].kind == GO_REG) s->has_reg = 1;
    s->cold = 1;
    /* topological + single-driver check */
    s->topo = 1;
    uint8_t *ready = (uint8_t *)calloc(W ? W : 1, 1);
    for (uint32_t w = 0; w < W; w++) {
        uint32_t nd = s->drivers_off[w + 1] - s->drivers_off[w];
        if (nd == 0) ready[w] = 1;
        if (nd > 1) { s->topo = 0; s->multi = 1; }
    }
    for (uint32_t c = 0; c < C && s->topo; c++) {
        const comp_t *cp = &s->comps[c];
        for (uint32_t i = 0; i < cp->n_inputs; i++) if (!ready[s->inputs[cp->in_off + i]]) s->topo = 0;
        if (cp->kind == GO_MUX && !ready[cp->select]) s->topo = 0;
        if (cp->kind == GO_BUF || cp->kind == GO_REG) s->topo = 0;  /* transient conflicts, Z outputs, state: unit-delay model only */
        ready[cp->output] = 1;
    }
    free(ready);
    return s;
}

void bs_free(bs_sim *s) {
    if (!s) return;
    free(s->comps); free(s->inputs); free(s->drivers_off); free(s->drivers);
    free(s->dv); free(s->dm); free(s->av); free(s->am); free(s->bv); free(s->bm); free(s->ov); free(s->om);
    free(s->rdv); free(s->rdm); free(s->rpv); free(s->rpm); free(s->ndv); free(s->ndm); free(s->npv); free(s->npm); free(s);
}

int bs_is_topological(const bs_sim *s) { return s->topo; }
uint64_t bs_last_applications(const bs_sim *s) { return s->applications; }

void bs_set_drive(bs_sim *s, uint32_t wire, const uint64_t *value, const uint64_t *valid) {
    memcpy(s->dv + (size_t)wire * s->words, value, (size_t)s->words * 8);
    memcpy(s->dm + (size_t)wire * s->words, valid, (size_t)s->words * 8);
}

void bs_get(const bs_sim *s, uint32_t wire, uint64_t *value, uint64_t *valid) {
    memcpy(value, s->av + (size_t)wire * s->words, (size_t)s->words * 8);
    memcpy(valid, s->am + (size_t)wire * s->words, (size_t)s->words * 8);
}

void bs_get_all(const bs_sim *s, uint64_t *value, uint64_t *valid) {
    memcpy(value, s->av, (size_t)s->n_wires * s->words * 8);
    memcpy(valid, s->am, (size_t)s->n_wires * s->words * 8);
}

/* F_c(W) for one 64-lane word k, reading planes (wv, wm). */
static inline void eval_comp(const bs_sim *s, const comp_t *cp, const uint64_t *wv, const uint64_t *wm, uint32_t k, uint64_t *ov, uint64_t *om) {
    const uint32_t ci = (uint32_t)(cp - s->comps);
    const uint32_t *in = s->inputs + cp->in_off;
    uint32_t T = s->words;
    uint64_t v, m;
    if (cp->kind == GO_MUX) {
        uint32_t nsel = 0; while ((1u << nsel) < cp->n_inputs) nsel++;
        /* 1-bit wires only => select is 1 bit => exactly 2 data inputs */
        (void)nsel;
        uint64_t sv = wv[(size_t)cp->select * T + k], sm = wm[(size_t)cp->select * T + k];
        uint64_t d0v = wv[(size_t)in[0] * T + k], d0m = wm[(size_t)in[0] * T + k];
        uint64_t d1v = wv[(size_t)in[1] * T + k], d1m = wm[(size_t)in[1] * T + k];
        uint64_t pickm = (~sv & d0m) | (sv & d1m), pickv = (~sv & d0v) | (sv & d1v);
        m = sm & pickm;
        v = (sm & pickv) | ~m;
    } else if (cp->kind == GO_NOT) {
        uint64_t a = wv[(size_t)in[0] * T + k]; m = wm[(size_t)in[0] * T + k];
        v = ~a | ~m;
    } else if (cp->kind == GO_REG) {
        /* rising edge = previous clock a valid 0, clock now a valid 1; enable 1 loads (High-Z reads X), 0 keeps,
           Z/X makes the value Undefined.  Writes the next stored value / clock history (committed with state B). */
        size_t si = (size_t)ci * T + k;
        uint64_t dv = wv[(size_t)in[0] * T + k], dm = wm[(size_t)in[0] * T + k];
        uint64_t ev = wv[(size_t)in[1] * T + k], em = wm[(size_t)in[1] * T + k];
        uint64_t cv = wv[(size_t)in[2] * T + k], cm = wm[(size_t)in[2] * T + k];
        uint64_t edge = s->rpm[si] & ~s->rpv[si] & cm & cv;
        uint64_t take = edge & ev & em, keep = ~edge | (edge & ~ev & em), undef = edge & ~em;
        m = (take & dm) | (keep & s->rdm[si]);
        v = (take & (dv | ~dm)) | (keep & s->rdv[si]) | undef;
        s->ndv[si] = v; s->ndm[si] = m; s->npv[si] = cv; s->npm[si] = cm;
    } else if (cp->kind == GO_ADD || cp->kind == GO_SUB) {
        /* 1-bit wires: sum and difference modulo 2 are both XOR; any invalid input bit makes the output Undefined */
        uint64_t a = wv[(size_t)in[0] * T + k], b2 = wv[(size_t)in[1] * T + k];
        m = wm[(size_t)in[0] * T + k] & wm[(size_t)in[1] * T + k];
        v = (a ^ b2) | ~m;
    } else if (cp->kind == GO_HGATE) {
        /* 1-bit wires: a reduction over one bit is the bit itself (High-Z read as X), inverted for NAND/NOR/XNOR */
        uint64_t a = wv[(size_t)in[0] * T + k]; m = wm[(size_t)in[0] * T + k];
        v = (cp->select >= GO_NAND ? ~a : a) | ~m;
    } else if (cp->kind == GO_SLICE || cp->kind == GO_MERGE) {
        /* 1-bit wires: a slice at offset 0 or a merge of one input is a copy; High-Z reads X */
        uint64_t a = wv[(size_t)in[0] * T + k]; m = wm[(size_t)in[0] * T + k];
        v = a | ~m;
    } else if (cp->kind == GO_BUF) {
        /* enable 1: the input, High-Z read as X; enable 0: High-Z (0,0); enable Z/X: X (1,0) */
        uint64_t a = wv[(size_t)in[0] * T + k], am = wm[(size_t)in[0] * T + k];
        uint64_t ev = wv[(size_t)in[1] * T + k], em = wm[(size_t)in[1] * T + k];
        uint64_t en1 = ev & em;
        m = en1 & am;
        v = (en1 & (a | ~am)) | ~em;
    } else {
        v = wv[(size_t)in[0] * T + k]; m = wm[(size_t)in[0] * T + k];
        for (uint32_t i = 1; i < cp->n_inputs; i++) {
            uint64_t bv = wv[(size_t)in[i] * T + k], bm = wm[(size_t)in[i] * T + k];
            switch (cp->kind) {
            case GO_AND: case GO_NAND: f_and(v, m, bv, bm, &v, &m); break;
            case GO_OR: case GO_NOR: f_or(v, m, bv, bm, &v, &m); break;
            default: f_xor(v, m, bv, bm, &v, &m); break;
            }
        }
        if (cp->kind >= GO_NAND) v = ~v | ~m;
    }
    *ov = v; *om = m;
}

/* One application B = G(A) on word k. Returns changed mask; *conflict gets the conflict mask. */
static uint64_t apply_word(bs_sim *s, uint32_t k, uint64_t *conflict) {
    uint32_t T = s->words;
    uint64_t changed = 0, conf = 0;
    for (uint32_t w = 0; w < s->n_wires; w++) {
        size_t idx = (size_t)w * T + k;
        uint64_t rv = s->dv[idx], rm = s->dm[idx];
        for (uint32_t d = s->drivers_off[w]; d < s->drivers_off[w + 1]; d++) {
            uint64_t ov, om;
            const uint32_t c = s->drivers[d];
            eval_comp(s, &s->comps[c], s->av, s->am, k, &ov, &om);
            conf |= (rv | rm) & (ov | om);
            rv |= ov; rm |= om;
            if (s->multi) { /* output-change check (differs from the wire check only with several drivers) */
                size_t oi = (size_t)c * T + k;
                changed |= (ov ^ s->ov[oi]) | (om ^ s->om[oi]);
                s->ov[oi] = ov; s->om[oi] = om;
            }
        }
        changed |= (rv ^ s->av[idx]) | (rm ^ s->am[idx]);
        s->bv[idx] = rv; s->bm[idx] = rm;
    }
    *conflict = conf;
    return changed;
}

/*
 * run_sim on every lane.  Outputs (each `words` long, bit l%64 of word l/64 = lane l):
 *   conflict[]  lanes that returned Err(conflict)
 *   unsettled[] lanes that returned MaxStepsReached (never also in conflict[])
 * State after a conflict is outside the parity contract.
 */
void bs_run(bs_sim *s, uint64_t max_steps, uint64_t *conflict, uint64_t *unsettled) {
    uint32_t T = s->words;
    uint64_t N = max_steps == UINT64_MAX ? UINT64_MAX : max_steps + 1;
    uint64_t last = N >= UINT64_MAX - 2 ? UINT64_MAX : N + 2; /* index of the final, check-only application */
    memset(conflict, 0, (size_t)T * 8); memset(unsettled, 0, (size_t)T * 8);
    size_t n = (size_t)s->n_wires * T;
    uint64_t k = 0;
    s->applications = 0;
    if (s->cold) {
        memcpy(s->av, s->dv, n * 8); memcpy(s->am, s->dm, n * 8);
        s->cold = 0; k = 1;
        if (s->n_comps == 0) return;
    }
    uint64_t *live = (uint64_t *)malloc((size_t)(T ? T : 1) * 8); /* lanes still running */
    for (uint32_t i = 0; i < T; i++) live[i] = ~0ull;
    for (;;) {
        k++;
        s->applications++;
        int any = 0;
        for (uint32_t w = 0; w < T; w++) {
            uint64_t conf, ch = apply_word(s, w, &conf);
            if (k < last) conflict[w] |= conf & live[w];
            live[w] &= ch;           /* a lane with no change has reached its fixed point */
            live[w] &= ~conflict[w]; /* gsim aborts the run on a conflict */
            if (k == last) unsettled[w] = live[w];
            if (live[w]) any = 1;
        }
        if (k == last) {
            /* check-only application: MaxStepsReached lanes keep W_{N+1} = A; settled lanes have A == B */
            break;
        }
        /* Commit B for lanes that are still running or just settled (A == B there anyway);
           conflicted lanes also take B (unspecified by contract). */
        uint64_t *t;
        t = s->av; s->av = s->bv; s->bv = t;
        t = s->am; s->am = s->bm; s->bm = t;
        if (s->has_reg) {
            t = s->rdv; s->rdv = s->ndv; s->ndv = t;  t = s->rdm; s->rdm = s->ndm; s->ndm = t;
            t = s->rpv; s->rpv = s->npv; s->npv = t;  t = s->rpm; s->rpm = s->npm; s->npm = t;
        }
        if (!any) break;
    }
    free(live);
}

/* DAG shortcut used for big netlists: evaluate components once in (topological) creation order.
   Valid only when bs_is_topological(); equals bs_run for any max_steps >= depth. */
int bs_eval_dag(bs_sim *s, uint64_t *conflict) {
    if (!s->topo) return -1;
    uint32_t T = s->words;
    memset(conflict, 0, (size_t)T * 8);
    for (uint32_t w = 0; w < s->n_wires; w++)
        if (s->drivers_off[w + 1] == s->drivers_off[w]) {
            memcpy(s->av + (size_t)w * T, s->dv + (size_t)w * T, (size_t)T * 8);
            memcpy(s->am + (size_t)w * T, s->dm + (size_t)w * T, (size_t)T * 8);
        }
    for (uint32_t c = 0; c < s->n_comps; c++) {
        const comp_t *cp = &s->comps[c];
        for (uint32_t k = 0; k < T; k++) {
            uint64_t ov, om;
            eval_comp(s, cp, s->av, s->am, k, &ov, &om);
            size_t idx = (size_t)cp->output * T + k;
            uint64_t dv = s->dv[idx], dm = s->dm[idx];
            conflict[k] |= (dv | dm) & (ov | om);
            s->av[idx] = ov | dv; s->am[idx] = om | dm;
        }
    }
    s->cold = 0;
    return 0;
}
