//! `B200Server`: a `SimServer` over the B200 engine's C ABI — written as the sibling of
//! `GsimServer` (crates/digilogic_gsim/src/lib.rs); the line numbers in the comments refer to it.
//! The same state machine is compiled and tested in C++ in the engine repository
//! (`b2_server_*`, digilogic_b200/csrc/capi.cpp).
use digilogic_netcode::*;
use std::num::NonZeroU8;

mod ffi;

struct Builder(*mut ffi::B2Builder);
struct Simulator(*mut ffi::B2Sim);

impl Drop for Builder {
    fn drop(&mut self) {
        unsafe { ffi::b2_builder_free(self.0) }
    }
}
impl Drop for Simulator {
    fn drop(&mut self) {
        unsafe { ffi::b2_sim_free(self.0) }
    }
}

// :5-14
enum ClientState {
    Building(Builder),
    Simulating(Simulator),
}

impl Default for ClientState {
    fn default() -> Self {
        Self::Building(Builder(unsafe { ffi::b2_builder_new() }))
    }
}

// :16-22
#[derive(Default)]
#[allow(missing_debug_implementations)]
pub struct B200Server {
    clients: ahash::AHashMap<ClientId, ClientState>,
    bit_plane_0: [u8; 32],
    bit_plane_1: [u8; 32],
}

// :55-66 (the C ABI returns AddComponentError ordinals 1..7)
fn component_error_to_server_error(rc: i32) -> ServerError {
    match rc {
        1 => ServerError::OutOfResources,    // TooManyComponents
        2 => ServerError::InvalidNetId,      // InvalidWireId
        3 => ServerError::WidthMismatch,     // WireWidthMismatch
        4 => ServerError::WidthIncompatible, // WireWidthIncompatible
        5 => ServerError::OutOfRange,        // OffsetOutOfRange
        6 | 7 => ServerError::InvalidInputCount, // TooFewInputs | InvalidInputCount
        _ => ServerError::Other,
    }
}

impl B200Server {
    // :33-38 (an unknown client id is an error here, a panic in GsimServer :26,30)
    fn builder(&mut self, client_id: ClientId) -> ServerResult<*mut ffi::B2Builder> {
        match self.clients.get_mut(&client_id).ok_or(ServerError::Other)? {
            ClientState::Building(builder) => Ok(builder.0),
            ClientState::Simulating(_) => Err(ServerError::InvalidState),
        }
    }

    // :40-52
    fn simulator(&mut self, client_id: ClientId) -> ServerResult<*mut ffi::B2Sim> {
        match self.clients.get_mut(&client_id).ok_or(ServerError::Other)? {
            ClientState::Building(_) => Err(ServerError::InvalidState),
            ClientState::Simulating(simulator) => Ok(simulator.0),
        }
    }

    fn check_output(builder: *mut ffi::B2Builder, width: NonZeroU8, output: u32) -> ServerResult<()> {
        let mut output_width = 0u8;
        if unsafe { ffi::b2_builder_wire_width(builder, output, &mut output_width) } != 0 {
            return Err(ServerError::InvalidNetId);
        }
        if width.get() != output_width {
            return Err(ServerError::WidthMismatch);
        }
        Ok(())
    }

    // gate_impl! :76-99
    fn gate(&mut self, client_id: ClientId, kind: i32, width: NonZeroU8, inputs: &[u32], output: u32) -> ServerResult<u32> {
        let builder = self.builder(client_id)?;
        Self::check_output(builder, width, output)?;
        let mut cell = 0u32;
        match unsafe { ffi::b2_add_gate(builder, kind, inputs.as_ptr(), inputs.len(), output, &mut cell) } {
            0 => Ok(cell),
            rc => Err(component_error_to_server_error(rc)),
        }
    }
}

impl SimServer for B200Server {
    type NetId = u32;
    type CellId = u32;

    fn max_clients(&mut self) -> usize {
        usize::MAX
    }

    fn client_connected(&mut self, client_id: ClientId) {
        self.clients.insert(client_id, ClientState::default());
    }

    fn client_disconnected(&mut self, client_id: ClientId) {
        self.clients.remove(&client_id);
    }

    // :117-121
    fn begin_build(&mut self, client_id: ClientId) -> ServerResult<()> {
        *self.clients.get_mut(&client_id).ok_or(ServerError::Other)? = ClientState::default();
        Ok(())
    }

    // :123-133
    fn end_build(&mut self, client_id: ClientId) -> ServerResult<()> {
        let client_state = self.clients.get_mut(&client_id).ok_or(ServerError::Other)?;
        match client_state {
            ClientState::Building(builder) => {
                let simulator = unsafe { ffi::b2_build(builder.0) }; // leaves the builder empty, like mem::take
                if simulator.is_null() {
                    return Err(ServerError::OutOfResources);
                }
                *client_state = ClientState::Simulating(Simulator(simulator));
                Ok(())
            }
            ClientState::Simulating(_) => Err(ServerError::InvalidState),
        }
    }

    // :135-139
    fn add_net(&mut self, client_id: ClientId, width: NonZeroU8) -> ServerResult<Self::NetId> {
        let builder = self.builder(client_id)?;
        let mut wire = 0u32;
        match unsafe { ffi::b2_add_wire(builder, width.get(), &mut wire) } {
            0 => Ok(wire),
            _ => Err(ServerError::OutOfResources),
        }
    }

    fn add_and_gate(&mut self, id: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(id, ffi::B2_AND, w, i, o)
    }
    fn add_or_gate(&mut self, id: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(id, ffi::B2_OR, w, i, o)
    }
    fn add_xor_gate(&mut self, id: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(id, ffi::B2_XOR, w, i, o)
    }
    fn add_nand_gate(&mut self, id: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(id, ffi::B2_NAND, w, i, o)
    }
    fn add_nor_gate(&mut self, id: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(id, ffi::B2_NOR, w, i, o)
    }
    fn add_xnor_gate(&mut self, id: ClientId, w: NonZeroU8, i: &[u32], o: u32) -> ServerResult<u32> {
        self.gate(id, ffi::B2_XNOR, w, i, o)
    }

    // :148-167
    fn add_not_gate(&mut self, client_id: ClientId, width: NonZeroU8, input: u32, output: u32) -> ServerResult<u32> {
        let builder = self.builder(client_id)?;
        Self::check_output(builder, width, output)?;
        let mut cell = 0u32;
        match unsafe { ffi::b2_add_not(builder, input, output, &mut cell) } {
            0 => Ok(cell),
            rc => Err(component_error_to_server_error(rc)),
        }
    }

    // :169-192 — inputs[0] is the select net
    fn add_mux(&mut self, client_id: ClientId, width: NonZeroU8, inputs: &[u32], output: u32) -> ServerResult<u32> {
        let builder = self.builder(client_id)?;
        Self::check_output(builder, width, output)?;
        let (select, inputs) = inputs.split_first().ok_or(ServerError::InvalidInputCount)?;
        let mut cell = 0u32;
        match unsafe { ffi::b2_add_mux(builder, inputs.as_ptr(), inputs.len(), *select, output, &mut cell) } {
            0 => Ok(cell),
            rc => Err(component_error_to_server_error(rc)),
        }
    }

    // :194-214
    fn set_net_drive(&mut self, client_id: ClientId, net: u32, bit_plane_0: &[u8], bit_plane_1: &[u8]) -> ServerResult<()> {
        let simulator = self.simulator(client_id)?;
        if bit_plane_0.len() > 32 || bit_plane_1.len() > 32 {
            return Err(ServerError::OutOfRange); // GsimServer panics on the slice index (:205-206)
        }
        let mut words0 = [0u32; 8];
        let mut words1 = [0u32; 8];
        bytemuck::cast_slice_mut(&mut words0)[..bit_plane_0.len()].copy_from_slice(bit_plane_0);
        bytemuck::cast_slice_mut(&mut words1)[..bit_plane_1.len()].copy_from_slice(bit_plane_1);
        let word_count = bit_plane_0.len().min(bit_plane_1.len()).div_ceil(4);
        match unsafe { ffi::b2_set_wire_drive(simulator, net, words0.as_ptr(), words1.as_ptr(), word_count) } {
            0 => Ok(()),
            _ => Err(ServerError::InvalidNetId),
        }
    }

    // :216-219 with simulation_result_to_server_result :68-74
    fn eval(&mut self, client_id: ClientId, max_steps: u64) -> ServerResult<()> {
        match unsafe { ffi::b2_run_sim(self.simulator(client_id)?, max_steps) } {
            0 => Ok(()),
            1 => Err(ServerError::MaxStepsReached),
            2 => Err(ServerError::DriverConflict),
            _ => Err(ServerError::Other),
        }
    }

    // :221-238
    fn get_net_state(&mut self, client_id: ClientId, net: u32) -> ServerResult<(NonZeroU8, &[u8], &[u8])> {
        let simulator = self.simulator(client_id)?;
        let mut width = 0u8;
        if unsafe { ffi::b2_wire_width(simulator, net, &mut width) } != 0 {
            return Err(ServerError::InvalidNetId); // GsimServer unwrap()s (:228)
        }
        let mut words0 = [0u32; 8];
        let mut words1 = [0u32; 8];
        if unsafe { ffi::b2_get_wire_state(simulator, net, words0.as_mut_ptr(), words1.as_mut_ptr(), 8) } != 0 {
            return Err(ServerError::Other);
        }
        self.bit_plane_0.copy_from_slice(bytemuck::cast_slice(&words0));
        self.bit_plane_1.copy_from_slice(bytemuck::cast_slice(&words1));
        let width = NonZeroU8::new(width).ok_or(ServerError::Other)?;
        Ok((width, &self.bit_plane_0, &self.bit_plane_1))
    }
}
