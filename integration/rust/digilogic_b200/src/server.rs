//! `B200Server` implements `digilogic_netcode::SimServer` over the C ABI of libdigilogic_b200.
//!
//! Behaviour follows the gsim-backed server it stands beside (crates/digilogic_gsim/src/lib.rs; the
//! `:NN` notes give the lines whose behaviour each piece reproduces): one netlist under construction
//! or one running simulator per client, the width check before every gate, select-first mux inputs,
//! plane bytes widened to u32 words, two 32-byte scratch planes for read-back.  Where that server
//! panics (unknown client :26,:30; more than 32 plane bytes :205; unknown net :228) this one
//! returns a `ServerError`, because nothing may unwind across the FFI boundary.
use crate::ffi;
use digilogic_netcode::*;
use std::collections::HashMap;
use std::num::NonZeroU8;
use std::os::raw::c_int;

/// Owning handle of a `b2_builder`.
struct Netlist(*mut ffi::B2Builder);
/// Owning handle of a `b2_sim`.
struct Engine(*mut ffi::B2Sim);

impl Drop for Netlist {
    fn drop(&mut self) {
        unsafe { ffi::b2_builder_free(self.0) }
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::b2_sim_free(self.0) }
    }
}

/// What a client currently owns (:5-14).
enum Phase {
    Assembling(Netlist),
    Running(Engine),
}

impl Phase {
    fn fresh() -> Self {
        Phase::Assembling(Netlist(unsafe { ffi::b2_builder_new() }))
    }
}

#[allow(missing_debug_implementations)]
pub struct B200Server {
    phases: HashMap<ClientId, Phase>,
    scratch: [[u8; 32]; 2],
}

impl Default for B200Server {
    fn default() -> Self {
        Self { phases: HashMap::new(), scratch: [[0; 32]; 2] }
    }
}

/// `b2_status` 1..7 are gsim's `AddComponentError` variants; map them like :55-66.
fn build_error(code: c_int) -> ServerError {
    match code {
        1 => ServerError::OutOfResources,
        2 => ServerError::InvalidNetId,
        3 => ServerError::WidthMismatch,
        4 => ServerError::WidthIncompatible,
        5 => ServerError::OutOfRange,
        6 | 7 => ServerError::InvalidInputCount,
        _ => ServerError::Other,
    }
}

fn cell_or_error(code: c_int, cell: u32) -> ServerResult<u32> {
    if code == 0 {
        Ok(cell)
    } else {
        Err(build_error(code))
    }
}

impl B200Server {
    fn netlist(&mut self, client: ClientId) -> ServerResult<*mut ffi::B2Builder> {
        match self.phases.get_mut(&client) {
            Some(Phase::Assembling(n)) => Ok(n.0),
            Some(Phase::Running(_)) => Err(ServerError::InvalidState),
            None => Err(ServerError::Other),
        }
    }

    fn engine(&mut self, client: ClientId) -> ServerResult<*mut ffi::B2Sim> {
        match self.phases.get_mut(&client) {
            Some(Phase::Running(e)) => Ok(e.0),
            Some(Phase::Assembling(_)) => Err(ServerError::InvalidState),
            None => Err(ServerError::Other),
        }
    }

    /// The check every `add_*` does first (:87-92): the declared width is the output net's width.
    fn netlist_for_output(&mut self, client: ClientId, width: NonZeroU8, output: u32) -> ServerResult<*mut ffi::B2Builder> {
        let netlist = self.netlist(client)?;
        let mut actual = 0u8;
        if unsafe { ffi::b2_builder_wire_width(netlist, output, &mut actual) } != 0 {
            return Err(ServerError::InvalidNetId);
        }
        if actual != width.get() {
            return Err(ServerError::WidthMismatch);
        }
        Ok(netlist)
    }

    fn gate(&mut self, client: ClientId, kind: c_int, width: NonZeroU8, inputs: &[u32], output: u32) -> ServerResult<u32> {
        let netlist = self.netlist_for_output(client, width, output)?;
        let mut cell = 0u32;
        let code = unsafe { ffi::b2_add_gate(netlist, kind, inputs.as_ptr(), inputs.len(), output, &mut cell) };
        cell_or_error(code, cell)
    }
}

/// Little-endian plane bytes -> u32 words, missing bytes zero (:199-206).
fn plane_words(bytes: &[u8]) -> [u32; 8] {
    let mut words = [0u32; 8];
    for (i, b) in bytes.iter().enumerate() {
        words[i / 4] |= (*b as u32) << (8 * (i % 4));
    }
    words
}

impl SimServer for B200Server {
    type NetId = u32;
    type CellId = u32;

    fn max_clients(&mut self) -> usize {
        usize::MAX
    }

    fn client_connected(&mut self, client: ClientId) {
        self.phases.insert(client, Phase::fresh());
    }

    fn client_disconnected(&mut self, client: ClientId) {
        self.phases.remove(&client);
    }

    fn begin_build(&mut self, client: ClientId) -> ServerResult<()> {
        let phase = self.phases.get_mut(&client).ok_or(ServerError::Other)?;
        *phase = Phase::fresh(); // always a new netlist (:117-121)
        Ok(())
    }

    fn end_build(&mut self, client: ClientId) -> ServerResult<()> {
        let netlist = self.netlist(client)?; // InvalidState when already running (:131)
        let sim = unsafe { ffi::b2_build(netlist) }; // compiles the netlist, empties the builder
        if sim.is_null() {
            return Err(ServerError::OutOfResources);
        }
        self.phases.insert(client, Phase::Running(Engine(sim)));
        Ok(())
    }

    fn add_net(&mut self, client: ClientId, width: NonZeroU8) -> ServerResult<u32> {
        let netlist = self.netlist(client)?;
        let mut net = 0u32;
        match unsafe { ffi::b2_add_wire(netlist, width.get(), &mut net) } {
            0 => Ok(net),
            _ => Err(ServerError::OutOfResources),
        }
    }

    fn add_and_gate(&mut self, c: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(c, ffi::B2_AND, w, i, o)
    }
    fn add_or_gate(&mut self, c: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(c, ffi::B2_OR, w, i, o)
    }
    fn add_xor_gate(&mut self, c: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(c, ffi::B2_XOR, w, i, o)
    }
    fn add_nand_gate(&mut self, c: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(c, ffi::B2_NAND, w, i, o)
    }
    fn add_nor_gate(&mut self, c: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(c, ffi::B2_NOR, w, i, o)
    }
    fn add_xnor_gate(&mut self, c: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(c, ffi::B2_XNOR, w, i, o)
    }

    fn add_not_gate(&mut self, client: ClientId, width: NonZeroU8, input: u32, output: u32) -> ServerResult<u32> {
        let netlist = self.netlist_for_output(client, width, output)?;
        let mut cell = 0u32;
        let code = unsafe { ffi::b2_add_not(netlist, input, output, &mut cell) };
        cell_or_error(code, cell)
    }

    /// `inputs[0]` is the select net, the rest are the data nets (:185-187).
    fn add_mux(&mut self, client: ClientId, width: NonZeroU8, inputs: &[u32], output: u32) -> ServerResult<u32> {
        let netlist = self.netlist_for_output(client, width, output)?;
        let Some((select, data)) = inputs.split_first() else {
            return Err(ServerError::InvalidInputCount);
        };
        let mut cell = 0u32;
        let code = unsafe { ffi::b2_add_mux(netlist, data.as_ptr(), data.len(), *select, output, &mut cell) };
        cell_or_error(code, cell)
    }

    fn set_net_drive(&mut self, client: ClientId, net: u32, plane_0: &[u8], plane_1: &[u8]) -> ServerResult<()> {
        let engine = self.engine(client)?;
        if plane_0.len() > 32 || plane_1.len() > 32 {
            return Err(ServerError::OutOfRange);
        }
        let (value, valid) = (plane_words(plane_0), plane_words(plane_1));
        let words = plane_0.len().min(plane_1.len()).div_ceil(4); // bits past the shorter plane are High-Z (:208)
        match unsafe { ffi::b2_set_wire_drive(engine, net, value.as_ptr(), valid.as_ptr(), words) } {
            0 => Ok(()),
            _ => Err(ServerError::InvalidNetId),
        }
    }

    fn eval(&mut self, client: ClientId, max_steps: u64) -> ServerResult<()> {
        let engine = self.engine(client)?;
        match unsafe { ffi::b2_run_sim(engine, max_steps) } {
            0 => Ok(()),                               // SimulationRunResult::Ok
            1 => Err(ServerError::MaxStepsReached),    // ::MaxStepsReached
            2 => Err(ServerError::DriverConflict),     // ::Err(_)
            _ => Err(ServerError::Other),              // engine failure (no device, CUDA error)
        }
    }

    fn get_net_state(&mut self, client: ClientId, net: u32) -> ServerResult<(NonZeroU8, &[u8], &[u8])> {
        let engine = self.engine(client)?;
        let mut width = 0u8;
        if unsafe { ffi::b2_wire_width(engine, net, &mut width) } != 0 {
            return Err(ServerError::InvalidNetId);
        }
        let (mut value, mut valid) = ([0u32; 8], [0u32; 8]);
        if unsafe { ffi::b2_get_wire_state(engine, net, value.as_mut_ptr(), valid.as_mut_ptr(), 8) } != 0 {
            return Err(ServerError::Other);
        }
        for i in 0..8 {
            self.scratch[0][4 * i..4 * i + 4].copy_from_slice(&value[i].to_le_bytes());
            self.scratch[1][4 * i..4 * i + 4].copy_from_slice(&valid[i].to_le_bytes());
        }
        let width = NonZeroU8::new(width).ok_or(ServerError::Other)?;
        Ok((width, &self.scratch[0], &self.scratch[1]))
    }
}
