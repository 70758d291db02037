//! Bindings of include/digilogic_b200.h — the entry points GsimServer's gsim calls map to
//! (crates/digilogic_gsim/src/lib.rs line numbers in the comments).
use std::os::raw::c_int;

#[repr(C)]
pub struct B2Builder {
    _p: [u8; 0],
}
#[repr(C)]
pub struct B2Sim {
    _p: [u8; 0],
}

#[repr(C)]
pub struct B2Batch {
    _p: [u8; 0],
}
#[repr(C)]
pub struct B2Comm {
    _p: [u8; 0],
}

/// b2_batch_info of include/digilogic_b200.h
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct B2BatchInfo {
    pub tile_words: u32,
    pub n_slots: u32,
    pub sweep: u32,
    pub tasks_per_pass: u32,
    pub grid_ctas: u32,
    pub gates_per_task: u32,
    pub n_ranks: u32,
    pub rank: u32,
    pub peers_mapped: u32,
    pub pair_launches: u32,
    pub store_bytes: u64,
    pub words_local: u64,
    pub last_tiles: u64,
    pub last_tickets: u64,
}

// b2_builder_option: the points of gsim's behaviour reconstructed without its source
pub const B2_BOPT_MUX_Z_PASSTHROUGH: c_int = 1;
pub const B2_BOPT_INITIAL_UNDEFINED: c_int = 2;
pub const B2_BOPT_BUFFER_Z_PASSTHROUGH: c_int = 3;
pub const B2_BOPT_COPY_Z_PASSTHROUGH: c_int = 4;
// b2_option (simulator): the switches a host is most likely to touch; the full list is in include/digilogic_b200.h
pub const B2_OPT_STEP_BUDGET_INCLUSIVE: c_int = 6;
pub const B2_OPT_CONFLICT_NEEDS_DISAGREE: c_int = 8;
pub const B2_OPT_PAIR_FUSION: c_int = 12; // 0 off, 1 on, 2 automatic (default)
// b2_batch_option
pub const B2_BATCH_OPT_SWEEP: c_int = 1;
pub const B2_BATCH_OPT_AUTO_GATHER: c_int = 5;
pub const B2_BATCH_OPT_DISCARD: c_int = 6;

pub const B2_AND: c_int = 0;
pub const B2_OR: c_int = 1;
pub const B2_XOR: c_int = 2;
pub const B2_NAND: c_int = 3;
pub const B2_NOR: c_int = 4;
pub const B2_XNOR: c_int = 5;

extern "C" {
    pub fn b2_builder_new() -> *mut B2Builder; // SimulatorBuilder::default()                 :12
    pub fn b2_builder_free(b: *mut B2Builder);
    pub fn b2_add_wire(b: *mut B2Builder, width: u8, out_wire: *mut u32) -> c_int; // add_wire  :135-139
    pub fn b2_builder_wire_width(b: *const B2Builder, wire: u32, out: *mut u8) -> c_int; // get_wire_width :87-92
    pub fn b2_add_gate(b: *mut B2Builder, kind: c_int, inputs: *const u32, n: usize, output: u32, out_cell: *mut u32) -> c_int; // add_*_gate :95
    pub fn b2_add_not(b: *mut B2Builder, input: u32, output: u32, out_cell: *mut u32) -> c_int; // add_not_gate :165
    pub fn b2_add_buffer(b: *mut B2Builder, input: u32, enable: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_buffer (no call site in digilogic_gsim yet)
    pub fn b2_add_add(b: *mut B2Builder, input_a: u32, input_b: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_add
    pub fn b2_add_sub(b: *mut B2Builder, input_a: u32, input_b: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_sub
    pub fn b2_add_mul(b: *mut B2Builder, input_a: u32, input_b: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_mul
    pub fn b2_add_compare(b: *mut B2Builder, op: c_int, input_a: u32, input_b: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_compare_* (op = B2_CMP_*)
    pub fn b2_add_shift(b: *mut B2Builder, kind: c_int, input: u32, amount: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_{left,logical_right,arithmetic_right}_shift
    pub fn b2_add_horizontal_gate(b: *mut B2Builder, kind: c_int, input: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_horizontal_*_gate
    pub fn b2_add_register(b: *mut B2Builder, data_in: u32, data_out: u32, enable: u32, clock: u32, out_cell: *mut u32) -> c_int; // gsim add_register
    pub fn b2_add_slice(b: *mut B2Builder, input: u32, offset: u32, output: u32, out_cell: *mut u32) -> c_int; // gsim add_slice
    pub fn b2_add_merge(b: *mut B2Builder, inputs: *const u32, n: usize, output: u32, out_cell: *mut u32) -> c_int; // gsim add_merge
    pub fn b2_add_mux(b: *mut B2Builder, inputs: *const u32, n: usize, select: u32, output: u32, out_cell: *mut u32) -> c_int; // add_multiplexer :190
    pub fn b2_build(b: *mut B2Builder) -> *mut B2Sim; // mem::take(builder).build()             :127
    pub fn b2_sim_free(s: *mut B2Sim);
    pub fn b2_wire_width(s: *const B2Sim, wire: u32, out: *mut u8) -> c_int; // Simulator::get_wire_width :228
    pub fn b2_set_wire_drive(s: *mut B2Sim, wire: u32, p0: *const u32, p1: *const u32, words: usize) -> c_int; // :209-213
    pub fn b2_run_sim(s: *mut B2Sim, max_steps: u64) -> c_int; // run_sim                       :218
    pub fn b2_get_wire_state(s: *mut B2Sim, wire: u32, p0: *mut u32, p1: *mut u32, words: usize) -> c_int; // :229-230
    pub fn b2_pack_sim_state(s: *mut B2Sim, wires: *const u32, n: u32, p0: *mut u8, p1: *mut u8, cap: usize, bit_len: *mut u64) -> c_int;
    pub fn b2_builder_set_option(b: *mut B2Builder, option: c_int, value: i64) -> c_int; // SURVEY A.6-4..7 switches
    pub fn b2_sim_clone(s: *const B2Sim) -> *mut B2Sim; // a second simulator of the same compiled netlist (shared device copy)
    pub fn b2_sim_set_option(s: *mut B2Sim, option: c_int, value: i64) -> c_int; // engine options (B2_OPT_*), copied into a batch's replicas
    pub fn b2_sim_plan_pairs(s: *const B2Sim, keep_wires: *const u32, n_keep: u32, out: *mut u64) -> c_int; // [u64; 6]: the level-pair plan, host only
    pub fn b2_sim_get_pair_info(s: *const B2Sim, launches: *mut u64, children: *mut u64, unstored: *mut u64) -> c_int;

    // ---- batch runner and multi-GPU gather (no gsim counterpart: L lanes = L Simulator::run_sim calls, lib.rs:216-219)
    pub fn b2_batch_new(sim: *const B2Sim, inputs: *const u32, n_inputs: u32, outputs: *const u32, n_outputs: u32, tile_words: u32, n_slots: u32, out: *mut *mut B2Batch) -> c_int;
    pub fn b2_batch_free(b: *mut B2Batch);
    pub fn b2_batch_set_device(b: *mut B2Batch, device: c_int) -> c_int;
    pub fn b2_batch_set_stream(b: *mut B2Batch, stream: *mut std::ffi::c_void) -> c_int;
    pub fn b2_batch_set_option(b: *mut B2Batch, option: c_int, value: i64) -> c_int;
    pub fn b2_batch_get_info(b: *mut B2Batch, out: *mut B2BatchInfo) -> c_int;
    pub fn b2_batch_run_generated(b: *mut B2Batch, seed: u64, mode: c_int, lane_word_begin: u64, n_words: u64, max_steps: u64, out_value: *mut u64, out_valid: *mut u64, out_stride_words: usize, out_ring_words: usize) -> c_int;
    pub fn b2_batch_run_planes(b: *mut B2Batch, in_value: *const u64, in_valid: *const u64, in_stride_words: usize, in_ring_words: usize, n_words: u64, max_steps: u64, out_value: *mut u64, out_valid: *mut u64, out_stride_words: usize, out_ring_words: usize) -> c_int;
    pub fn b2_batch_wait(b: *mut B2Batch, out_worst: *mut c_int) -> c_int;
    pub fn b2_comm_get_unique_id(id: *mut u8) -> c_int; // [u8; 128], ncclGetUniqueId
    pub fn b2_comm_init_rank(out: *mut *mut B2Comm, n_ranks: c_int, rank: c_int, id: *const u8, device: c_int) -> c_int; // ncclCommInitRank
    pub fn b2_comm_free(c: *mut B2Comm);
    pub fn b2_batch_bind_comm(b: *mut B2Batch, comm: *mut B2Comm, words_local: u64) -> c_int;
    pub fn b2_batch_export_gather_handle(b: *mut B2Batch, handle: *mut u8) -> c_int; // [u8; 128]: two cudaIpcMemHandle_t
    pub fn b2_batch_import_gather_handles(b: *mut B2Batch, handles: *const u8) -> c_int; // [n_ranks][128]
    pub fn b2_batch_gather_outputs(b: *mut B2Batch) -> c_int;
    pub fn b2_batch_gathered(b: *mut B2Batch, dev_value: *mut *mut u64, dev_valid: *mut *mut u64) -> c_int;
}
