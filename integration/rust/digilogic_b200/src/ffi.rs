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
}
