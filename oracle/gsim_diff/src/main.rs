//! Runs every case of a cases.json (written by make_cases.py) on real gsim 1.1.4 through exactly the
//! calls digilogic_gsim makes (crates/digilogic_gsim/src/lib.rs:12-230) and dumps, after every
//! run_sim, the result and the two bit planes of every wire.
use gsim::*;
use serde::{Deserialize, Serialize};
use std::num::NonZeroU8;

#[derive(Deserialize)]
struct Comp {
    kind: String, // and | or | xor | nand | nor | xnor | not | mux | buffer (inputs = [data, enable]) | slice (select = offset) | merge | register (inputs = [data_in, enable, clock]) | add | sub | hand..hxnor (horizontal gates)
    inputs: Vec<u32>,
    select: Option<u32>,
    output: u32,
}
#[derive(Deserialize)]
struct Run {
    drives: Vec<(u32, Vec<u32>, Vec<u32>)>, // (wire, plane0 words, plane1 words)
    max_steps: u64,
}
#[derive(Deserialize)]
struct Case {
    name: String,
    wires: Vec<u8>,
    comps: Vec<Comp>,
    runs: Vec<Run>,
}
#[derive(Serialize)]
struct RunOut {
    result: &'static str, // Ok | MaxStepsReached | Err
    wires: Vec<(Vec<u32>, Vec<u32>)>,
}
#[derive(Serialize)]
struct CaseOut {
    name: String,
    build_errors: Vec<Option<String>>,
    runs: Vec<RunOut>,
}

fn main() {
    let path = std::env::args().nth(1).expect("usage: gsim_diff cases.json");
    let cases: Vec<Case> = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    let mut out = Vec::new();
    for case in cases {
        let mut builder = SimulatorBuilder::default();
        let wires: Vec<WireId> = case.wires.iter().map(|w| builder.add_wire(NonZeroU8::new(*w).unwrap()).unwrap()).collect();
        let mut build_errors = Vec::new();
        for c in &case.comps {
            let ins: Vec<WireId> = c.inputs.iter().map(|i| wires[*i as usize]).collect();
            let o = wires[c.output as usize];
            let r = match c.kind.as_str() {
                "and" => builder.add_and_gate(&ins, o),
                "or" => builder.add_or_gate(&ins, o),
                "xor" => builder.add_xor_gate(&ins, o),
                "nand" => builder.add_nand_gate(&ins, o),
                "nor" => builder.add_nor_gate(&ins, o),
                "xnor" => builder.add_xnor_gate(&ins, o),
                "not" => builder.add_not_gate(ins[0], o),
                "mux" => builder.add_multiplexer(&ins, wires[c.select.unwrap() as usize], o),
                "buffer" => builder.add_buffer(ins[0], ins[1], o),
                "slice" => builder.add_slice(ins[0], c.select.unwrap() as u8, o),
                "merge" => builder.add_merge(&ins, o),
                "register" => builder.add_register(ins[0], o, ins[1], ins[2]),
                "add" => builder.add_add(ins[0], ins[1], o),
                "sub" => builder.add_sub(ins[0], ins[1], o),
                "mul" => builder.add_mul(ins[0], ins[1], o),
                "cmp_eq" => builder.add_compare_equal(ins[0], ins[1], o),
                "cmp_ne" => builder.add_compare_not_equal(ins[0], ins[1], o),
                "cmp_ltu" => builder.add_compare_less_than(ins[0], ins[1], o),
                "cmp_gtu" => builder.add_compare_greater_than(ins[0], ins[1], o),
                "cmp_leu" => builder.add_compare_less_than_or_equal(ins[0], ins[1], o),
                "cmp_geu" => builder.add_compare_greater_than_or_equal(ins[0], ins[1], o),
                "cmp_lts" => builder.add_compare_less_than_signed(ins[0], ins[1], o),
                "cmp_gts" => builder.add_compare_greater_than_signed(ins[0], ins[1], o),
                "cmp_les" => builder.add_compare_less_than_or_equal_signed(ins[0], ins[1], o),
                "cmp_ges" => builder.add_compare_greater_than_or_equal_signed(ins[0], ins[1], o),
                "shl" => builder.add_left_shift(ins[0], ins[1], o),
                "lshr" => builder.add_logical_right_shift(ins[0], ins[1], o),
                "ashr" => builder.add_arithmetic_right_shift(ins[0], ins[1], o),
                "hand" => builder.add_horizontal_and_gate(ins[0], o),
                "hor" => builder.add_horizontal_or_gate(ins[0], o),
                "hxor" => builder.add_horizontal_xor_gate(ins[0], o),
                "hnand" => builder.add_horizontal_nand_gate(ins[0], o),
                "hnor" => builder.add_horizontal_nor_gate(ins[0], o),
                "hxnor" => builder.add_horizontal_xnor_gate(ins[0], o),
                k => panic!("unknown kind {k}"),
            };
            build_errors.push(r.err().map(|e| format!("{e:?}")));
        }
        let mut sim = builder.build();
        let mut runs = Vec::new();
        for run in &case.runs {
            for (w, p0, p1) in &run.drives {
                let state = LogicState::from_bit_planes(p0, p1);
                sim.set_wire_drive(wires[*w as usize], &state).unwrap();
            }
            let result = match sim.run_sim(run.max_steps) {
                SimulationRunResult::Ok => "Ok",
                SimulationRunResult::MaxStepsReached => "MaxStepsReached",
                SimulationRunResult::Err(_) => "Err",
            };
            let mut planes = Vec::new();
            for w in &wires {
                let width = sim.get_wire_width(*w).unwrap();
                let (p0, p1) = sim.get_wire_state(*w).unwrap().to_bit_planes(width);
                planes.push((p0.to_vec(), p1.to_vec()));
            }
            runs.push(RunOut { result, wires: planes });
        }
        out.push(CaseOut { name: case.name, build_errors, runs });
    }
    println!("{}", serde_json::to_string(&out).unwrap());
}
