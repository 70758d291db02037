//! B200-native `SimServer` for digilogic: the engine behind `SimulationEngine::GsimCompute`.
//! See INTEGRATION.md in the engine repository.
mod ffi;
mod server;

pub use server::B200Server;
