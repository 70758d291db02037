/* gsim_oracle.h — shared types of the CPU oracle (TEST INFRASTRUCTURE ONLY; see gsim_oracle.c). */
#ifndef GSIM_ORACLE_H
#define GSIM_ORACLE_H
#include <stdint.h>

#define GO_MAX_ATOMS 8 /* width <= 255 bits => <= 8 u32 atoms (lib.rs:199-209) */

typedef struct { uint32_t state, valid; } atom_t;

enum { GO_AND = 0, GO_OR = 1, GO_XOR = 2, GO_NAND = 3, GO_NOR = 4, GO_XNOR = 5, GO_NOT = 6, GO_MUX = 7,
       GO_BUF = 8,  /* tri-state buffer: inputs[0] = data, inputs[1] = 1-bit enable */
       GO_SLICE = 9,  /* output = input[offset .. offset + width(output)); offset in `select` */
       GO_MERGE = 10, /* output = inputs concatenated, inputs[0] in the low bits */
       GO_REG = 11,   /* register: inputs = {data_in, enable (1 bit), clock (1 bit)}; loads data_in on a rising clock edge */
       GO_HGATE = 12, /* horizontal (reduction) gate: the 1-bit output is the AND/OR/.. (kind in `select`) of all bits of inputs[0] */
       GO_ADD = 13, GO_SUB = 14, /* output = inputs[0] +/- inputs[1] modulo 2^width; Undefined if any input bit is not valid */
       GO_MUL = 15,   /* output = inputs[0] * inputs[1] modulo 2^width; Undefined if any input bit is not valid */
       GO_CMP = 16,   /* 1-bit output = inputs[0] <op> inputs[1], op (GO_CMP_*) in `select`; Undefined if any input bit is not valid */
       GO_SHL = 17, GO_LSHR = 18, GO_ASHR = 19 /* output = inputs[0] shifted by the unsigned value of inputs[1] (zero / sign fill;
                         amounts >= width shift everything out); Undefined if any bit of either input is not valid */ };

enum { GO_CMP_EQ = 0, GO_CMP_NE = 1, GO_CMP_LTU = 2, GO_CMP_GTU = 3, GO_CMP_LEU = 4, GO_CMP_GEU = 5,
       GO_CMP_LTS = 6, GO_CMP_GTS = 7, GO_CMP_LES = 8, GO_CMP_GES = 9 };

/* AddComponentError ordinals, names pinned by crates/digilogic_gsim/src/lib.rs:55-66 */
enum {
    GO_OK = 0,
    GO_ERR_TOO_MANY_COMPONENTS = 1,
    GO_ERR_INVALID_WIRE_ID = 2,
    GO_ERR_WIRE_WIDTH_MISMATCH = 3,
    GO_ERR_WIRE_WIDTH_INCOMPATIBLE = 4,
    GO_ERR_OFFSET_OUT_OF_RANGE = 5,
    GO_ERR_TOO_FEW_INPUTS = 6,
    GO_ERR_INVALID_INPUT_COUNT = 7,
};

/* SimulationRunResult, crates/digilogic_gsim/src/lib.rs:68-74 */
enum { GO_RUN_OK = 0, GO_RUN_MAX_STEPS = 1, GO_RUN_CONFLICT = 2 };

/* SURVEY.md A.6 switches (defaults = the "believed" behaviour). */
typedef struct {
    int steps_check_ge;        /* A.6-1: 0 => `steps > max_steps` (default), 1 => `>=` */
    int conflict_needs_disagree; /* A.6-2: 0 => both non-Z conflicts (default) */
    int mux_z_passthrough;     /* A.6-4: 0 => Z on selected input reads X (default) */
    int initial_undefined;     /* A.6-5: 0 => High-Z initial state (default) */
    int buffer_z_passthrough;  /* A.6-6: 0 => an enabled buffer passes a High-Z input bit on as X (default), 1 => as Z */
    int copy_z_passthrough;    /* A.6-7: 0 => slices and merges pass a High-Z input bit on as X (default: only buffers output Z) */
} go_config_t;

typedef struct {
    uint8_t kind;
    uint32_t n_inputs;   /* data inputs (mux: 2^k data inputs, select separate) */
    uint32_t in_off;     /* into inputs[] */
    uint32_t select;     /* mux only */
    uint32_t output;
} comp_t;

typedef struct {
    /* wires */
    uint32_t n_wires, cap_wires;
    uint8_t *width;
    /* components */
    uint32_t n_comps, cap_comps;
    comp_t *comps;
    uint32_t n_inputs, cap_inputs;
    uint32_t *inputs;
} go_builder;

typedef struct {
    uint32_t n_wires, n_comps;
    uint8_t *width;
    uint8_t *natoms;
    comp_t *comps;
    uint32_t *inputs;
    /* graph: per wire, the components it drives ("driving") and its drivers */
    uint32_t *driving_off, *driving;   /* CSR, deduplicated */
    uint32_t *drivers_off, *drivers;   /* CSR */
    /* state: GO_MAX_ATOMS atoms per wire would waste memory on 1-bit nets, so
       atoms are packed: atom_off[w] .. atom_off[w]+natoms[w] */
    uint32_t *atom_off;
    uint32_t *catom_off;               /* per component output */
    uint32_t n_watoms, n_catoms;
    atom_t *drive, *state, *output;
    atom_t *rdata;                     /* registers: stored value, parallel to output[] (other components: unused) */
    int8_t *rprev;                     /* registers: clock at the previous update: -1 unknown, 0, 1 */
    /* queues */
    uint32_t *wq, *cq, *tmp;
    uint32_t n_wq, n_cq;
    uint64_t steps_done;               /* step_sim calls made by the last run_sim */
    uint64_t comp_evals;               /* component evaluations by the last run_sim */
    int owns_graph;                    /* clones share the immutable graph arrays of their source */
} go_sim;


go_builder *go_builder_new(void);
void go_builder_free(go_builder *b);
go_sim *go_build(const go_builder *b);
void go_sim_free(go_sim *s);
int go_run_sim(go_sim *s, uint64_t max_steps);

#endif
