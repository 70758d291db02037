"""CPU-side checks of the C ABI: every symbol the header declares is exported, the builder
validates like gsim (variants at reference crates/digilogic_gsim/src/lib.rs:55-66), the netlist
compiler's levelisation, the B200Server state machine (lib.rs:5-52,101-239) — and that the
product fails loudly, not quietly, without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import digilogic_b200 as d
from digilogic_b200 import _ffi, netgen
from digilogic_b200.gsim import AddComponentError, ComponentError
from digilogic_b200.server import B200Server, ServerError, ServerErrorException

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "digilogic_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    lib = C.CDLL(_ffi._build.LIB_PATH) if os.path.exists(_ffi._build.LIB_PATH) else _ffi.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    # and the ctypes table covers the header exactly
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    assert b"sm_100a" in _ffi.lib().b2_version()


def test_builder_validation_matches_gsim_variants():
    b = d.SimulatorBuilder()
    w1, w1b, w4 = b.add_wire(1), b.add_wire(1), b.add_wire(4)
    for fn, err in [
        (lambda: b.add_and_gate([w1, 99], w1b), AddComponentError.InvalidWireId),
        (lambda: b.add_and_gate([w1, w1b], 99), AddComponentError.InvalidWireId),
        (lambda: b.add_or_gate([w1], w1b), AddComponentError.TooFewInputs),
        (lambda: b.add_xor_gate([w1, w4], w1b), AddComponentError.WireWidthMismatch),
        (lambda: b.add_not_gate(w4, w1), AddComponentError.WireWidthMismatch),
        (lambda: b.add_multiplexer([w1, w1b, w1], w1, w1b), AddComponentError.InvalidInputCount),
        (lambda: b.add_multiplexer([w1, w1b], w4, w1b), AddComponentError.WireWidthIncompatible),
        (lambda: b.add_multiplexer([w1, w4], w1, w1b), AddComponentError.WireWidthMismatch),
        (lambda: b.add_buffer(w1, w1b, 99), AddComponentError.InvalidWireId),
        (lambda: b.add_buffer(w4, w1, w1b), AddComponentError.WireWidthMismatch),
        (lambda: b.add_buffer(w4, w4, w4), AddComponentError.WireWidthIncompatible),
    ]:
        with pytest.raises(ComponentError) as e:
            fn()
        assert e.value.error == err
    assert b.get_wire_width(w4) == 4
    assert b.add_and_gate([w1, w1b], w1b) == 0 and b.add_not_gate(w1, w1b) == 1  # ComponentIds count up


def test_same_validation_as_oracle():
    """The engine's builder and the oracle's builder accept and reject the same calls."""
    import oracle
    rng = np.random.default_rng(5)
    gb, ob = d.SimulatorBuilder(), oracle.Builder()
    widths = [1, 1, 1, 2, 2, 4, 8, 1]
    for w in widths:
        assert gb.add_wire(w) == ob.add_wire(w)
    for _ in range(1500):
        kind = int(rng.integers(0, 15))
        n = int(rng.integers(0, 5))
        ins = [int(x) for x in rng.integers(0, len(widths) + 1, n)]
        out = int(rng.integers(0, len(widths) + 1))
        sel = int(rng.integers(0, len(widths) + 1))
        x, y, off = (int(v) for v in rng.integers(0, len(widths) + 1, 3))
        def call(b):
            try:
                if kind == 14:
                    return ("ok", b.add_horizontal_gate(off % 6, x, out))
                if kind == 13:
                    return ("ok", b.add_sub(x, y, out))
                if kind == 12:
                    return ("ok", b.add_add(x, y, out))
                if kind == 11:
                    return ("ok", b.add_register(x, out, y, sel))
                if kind == 10:
                    return ("ok", b.add_merge(ins, out))
                if kind == 9:
                    return ("ok", b.add_slice(x, off, out))
                if kind == 8:
                    return ("ok", b.add_buffer(x, y, out))
                if kind == 7:
                    return ("ok", b.add_multiplexer(ins, sel, out))
                if kind == 6:
                    return ("ok", b.add_not_gate(ins[0] if ins else 0, out))
                return ("ok", b.add_gate(kind, ins, out))
            except ComponentError as e:
                return ("err", e.error.name)
            except oracle.OracleError as e:
                return ("err", str(e))
        assert call(gb) == call(ob)


def test_compile_info_levels():
    gen = netgen.random_dag(n_gates=4000, depth=40, n_inputs=64, n_outputs=16)
    sim = gen.emit(d.SimulatorBuilder()).build()
    i = sim.info
    assert (i.n_wires, i.n_rows, i.n_sources, i.n_gates2, i.n_gates_wide) == (4064, 4064, 64, 4000, 0)
    assert i.depth == 40 and i.cyclic == 0 and i.multi_driver == 0
    add = netgen.ripple_adder(4).emit(d.SimulatorBuilder()).build().info
    assert add.n_gates2 == 20 and add.depth == 9 and add.n_sources == 9
    mul = netgen.array_multiplier(32).emit(d.SimulatorBuilder()).build().info
    assert mul.n_gates2 == 5888 and mul.cyclic == 0
    b = d.SimulatorBuilder()
    s_, r_, q, qn, w8, o8 = b.add_wire(), b.add_wire(), b.add_wire(), b.add_wire(), b.add_wire(8), b.add_wire(8)
    b.add_nand_gate([s_, qn], q)
    b.add_nand_gate([r_, q], qn)
    b.add_not_gate(w8, o8)
    b.add_and_gate([s_, r_, q], qn)
    i = b.build().info
    assert i.cyclic == 1 and i.depth == 0 and i.multi_driver == 1
    assert i.n_sources == 10 and i.n_rows == 10 + 11 and i.n_gates2 == 2 + 8 and i.n_gates3 == 1 and i.n_gates_wide == 0


def test_compile_info_of_the_wider_component_set():
    """Row layout of tri-state nets, registers, slices, adders (host-only: b2_build needs no device)."""
    b = d.SimulatorBuilder()
    d0, d1, e0, e1, bus, nb = (b.add_wire() for _ in range(6))
    b.add_buffer(d0, e0, bus)
    b.add_buffer(d1, e1, bus)
    b.add_not_gate(bus, nb)
    i = b.build().info
    # 4 sources, 3 gate rows (each buffer keeps its own row) + 1 resolved row for the bus; unit-delay path only
    assert (i.n_sources, i.n_gates2, i.n_rows, i.cyclic, i.multi_driver) == (4, 3, 4 + 3 + 1, 1, 0)
    b = d.SimulatorBuilder()
    din, q, en, ck = b.add_wire(4), b.add_wire(4), b.add_wire(), b.add_wire()
    b.add_register(din, q, en, ck)
    i = b.build().info
    # 6 source bits; one hidden previous-clock gate (fast class) + 4 register bits (CSR class)
    assert (i.n_sources, i.n_gates2, i.n_gates_wide, i.n_rows, i.cyclic) == (6, 1, 4, 6 + 5, 1)
    b = d.SimulatorBuilder()
    x, y, sm, hi, par = b.add_wire(8), b.add_wire(8), b.add_wire(8), b.add_wire(4), b.add_wire(1)
    b.add_add(x, y, sm)
    b.add_slice(sm, 4, hi)
    b.add_horizontal_gate(d.gsim.XOR, hi, par)
    i = b.build().info
    # adder bits on the CSR class (8), slice bits are one-input fast gates (4), the 4-input reduction is a CSR gate;
    # acyclic: adder (1) -> slice (2) -> reduction (3)
    assert (i.n_sources, i.n_gates2, i.n_gates_wide, i.cyclic, i.depth) == (16, 4, 8 + 1, 0, 3)


def test_netlist_compiler_fuzz_under_sanitizers(tmp_path):
    """tests/netlist_fuzz.cpp: random builder calls over every component kind, compile(), invariants of the SoA arrays —
    host-only, built with AddressSanitizer + UBSan."""
    exe = str(tmp_path / "netlist_fuzz")
    res = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                          os.path.join(ROOT, "tests", "netlist_fuzz.cpp"), os.path.join(ROOT, "digilogic_b200", "csrc", "netlist.cpp")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe, "300"], capture_output=True, text=True)
    assert run.returncode == 0 and "netlist_fuzz ok" in run.stdout, run.stderr[-2000:]


def test_level_pair_planner_executed_on_a_scalar_model(tmp_path):
    """tests/pair_plan_check.cpp: the level-pair plan (B2_OPT_PAIR_FUSION, csrc/pair_plan.hpp) of random netlists, executed on
    a 4-state scalar model that tracks which rows hold data — every gate once, no read of a dropped or unstored row, kept
    rows equal to a plain evaluation; host-only, AddressSanitizer + UBSan."""
    exe = str(tmp_path / "pair_plan_check")
    res = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                          os.path.join(ROOT, "tests", "pair_plan_check.cpp"), os.path.join(ROOT, "digilogic_b200", "csrc", "netlist.cpp")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe, "300"], capture_output=True, text=True)
    assert run.returncode == 0 and "pair_plan_check ok" in run.stdout, run.stderr[-2000:]


def test_build_takes_the_builder():
    b = d.SimulatorBuilder()
    a = b.add_wire()
    sim = b.build()
    assert sim.info.n_wires == 1
    assert b.add_wire() == 0, "after build() the builder is an empty default builder (mem::take)"
    assert a == 0


def test_server_state_machine_and_errors():
    sv = B200Server()
    assert sv.max_clients() == 2**64 - 1
    with pytest.raises(ServerErrorException) as e:   # the reference panics on an unknown client (gsim/lib.rs:26)
        sv.begin_build(1)
    assert e.value.error == ServerError.Other
    sv.client_connected(1)
    sv.begin_build(1)
    a, b_, o = sv.add_net(1, 1), sv.add_net(1, 1), sv.add_net(1, 1)
    w4 = sv.add_net(1, 4)
    assert (a, b_, o, w4) == (0, 1, 2, 3)
    assert sv.add_and_gate(1, 1, [a, b_], o) == 0
    for fn, err in [
        (lambda: sv.add_or_gate(1, 4, [a, b_], o), ServerError.WidthMismatch),        # width != output width
        (lambda: sv.add_or_gate(1, 1, [a, b_], 77), ServerError.InvalidNetId),        # unknown output
        (lambda: sv.add_xor_gate(1, 1, [a, 77], o), ServerError.InvalidNetId),
        (lambda: sv.add_nand_gate(1, 1, [a], o), ServerError.InvalidInputCount),      # TooFewInputs
        (lambda: sv.add_nor_gate(1, 1, [a, w4], o), ServerError.WidthMismatch),
        (lambda: sv.add_not_gate(1, 1, w4, o), ServerError.WidthMismatch),
        (lambda: sv.add_mux(1, 1, [a, a, b_, o], o), ServerError.InvalidInputCount),  # 3 data inputs
        (lambda: sv.add_mux(1, 1, [w4, a, b_], o), ServerError.WidthIncompatible),    # select width
        (lambda: sv.set_net_drive(1, a, b"\1", b"\1"), ServerError.InvalidState),     # still Building
        (lambda: sv.eval(1, 10), ServerError.InvalidState),
        (lambda: sv.get_net_state(1, a), ServerError.InvalidState),
    ]:
        with pytest.raises(ServerErrorException) as e:
            fn()
        assert e.value.error == err, (err, e.value.error)
    assert sv.add_mux(1, 1, [a, b_, o], o) == 1          # inputs[0] is the select
    sv.end_build(1)
    for fn in (lambda: sv.end_build(1), lambda: sv.add_net(1, 1), lambda: sv.add_and_gate(1, 1, [a, b_], o)):
        with pytest.raises(ServerErrorException) as e:
            fn()
        assert e.value.error == ServerError.InvalidState
    with pytest.raises(ServerErrorException) as e:       # > 32 plane bytes panics in the reference (gsim/lib.rs:205)
        sv.set_net_drive(1, a, bytes(33), bytes(33))
    assert e.value.error == ServerError.OutOfRange
    sv.begin_build(1)                                    # always resets to a fresh builder (gsim/lib.rs:117-121)
    assert sv.add_net(1, 1) == 0
    sv.client_connected(1)                               # Adapter::client_disconnected calls this (server.rs:296-299)
    assert sv.add_net(1, 1) == 0
    sv.client_disconnected(1)
    with pytest.raises(ServerErrorException):
        sv.add_net(1, 1)


def test_no_cpu_fallback():
    """Without a CUDA device the engine must refuse, and nothing in the product imports the oracle."""
    for root, _, files in os.walk(os.path.join(ROOT, "digilogic_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "gsim_oracle" not in src and "bitsliced.c" not in src, f
    if _has_gpu():
        pytest.skip("GPU present")
    b = d.SimulatorBuilder()
    a, o = b.add_wire(), b.add_wire()
    b.add_not_gate(a, o)
    sim = b.build()
    for fn in (lambda: sim.run_sim(10), lambda: sim.set_lanes(64), lambda: sim.set_wire_drive(a, d.LogicState(1, 1)),
               lambda: sim.get_wire_state(o)):
        with pytest.raises(d.B200Error) as e:
            fn()
        assert e.value.code == 18 and "no CPU path" in str(e.value)


def _compile_c_client(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_abi_smoke")
    lib_dir = os.path.dirname(_ffi._build.LIB_PATH)
    _ffi.lib()
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", lib_dir, "-ldigilogic_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    return exe


def test_header_is_c99_and_a_c_client_links(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and a plain-C program links against the library.
    Without a GPU the client must stop at B2_ERR_NO_DEVICE (exit 77), never compute on the CPU."""
    import subprocess
    exe = _compile_c_client(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True)
    if _has_gpu():
        assert res.returncode == 0, res.stderr
    else:
        assert res.returncode == 77 and "no CUDA device" in res.stderr, (res.returncode, res.stderr)


@pytest.mark.gpu
def test_c_client_on_gpu(tmp_path):
    import subprocess
    res = subprocess.run([_compile_c_client(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0 and "c_abi_smoke ok" in res.stdout, res.stderr


def test_auto_tile_words():
    from digilogic_b200.batch import auto_tile_words
    cfg3 = netgen.random_dag(n_gates=200_000, depth=40, n_inputs=64, n_outputs=16).emit(d.SimulatorBuilder()).build()   # 5000 per level
    assert auto_tile_words(cfg3, 262144) == 256
    cfg5 = netgen.random_dag(n_gates=500_000, depth=10, n_inputs=64, n_outputs=16).emit(d.SimulatorBuilder()).build()   # 50 000 per level
    assert auto_tile_words(cfg5, 131072) == 16 or auto_tile_words(cfg5, 131072) == 32
    mul = netgen.array_multiplier(32).emit(d.SimulatorBuilder()).build()
    assert auto_tile_words(mul, 16384) == 4096
    assert auto_tile_words(mul, 96) == 32 and auto_tile_words(mul, 7) == 1


def test_batch_and_clone_host_side_without_a_device():
    """b2_sim_clone, b2_batch_new and the option setters are host-only; running needs a device (no CPU path)."""
    from digilogic_b200.batch import Batch, auto_batch_shape
    from digilogic_b200.gsim import B200Error, ComponentError
    gen = netgen.random_dag(n_gates=200_000, depth=40, n_inputs=64, n_outputs=16)          # 5000 gates per level, like config 3
    sim = gen.emit(d.SimulatorBuilder()).build()
    assert auto_batch_shape(sim, 262144) == (128, 3)
    c = sim.clone()
    assert c.info.n_rows == sim.info.n_rows and c.info.depth == sim.info.depth == 40
    b = Batch(sim, gen.inputs, gen.outputs, 128, 3)
    b.set_option(Batch.OPT_SWEEP, 1)
    b.set_option(Batch.OPT_DISCARD, 0)
    assert b.info.sweep == 1 and b.info.tile_words == 128 and b.info.n_slots == 3
    with pytest.raises(B200Error):
        b.set_option(Batch.OPT_UNROLL, 3)
    with pytest.raises(ComponentError):
        Batch(sim, [gen.n_wires], gen.outputs, 128, 3)                                     # InvalidWireId
    if not _has_gpu():
        out = np.zeros((16, 128), np.uint64)
        with pytest.raises(B200Error) as e:
            b.run_generated(1, 0, 0, 128, 10_000, out, out.copy())
        assert e.value.code == 18                                                          # B2_ERR_NO_DEVICE
    wide = d.SimulatorBuilder()
    w4, o4 = wide.add_wire(4), wide.add_wire(4)
    wide.add_not_gate(w4, o4)
    with pytest.raises(ComponentError):
        Batch(wide.build(), [w4], [o4], 32, 1)                                             # lane planes carry 1-bit wires


def test_builder_switches_survive_build():
    b = d.SimulatorBuilder()
    b.set_option(b.OPT_MUX_Z_PASSTHROUGH, 1)
    with pytest.raises(d.gsim.B200Error):
        b.set_option(99, 1)
    a, o = b.add_wire(), b.add_wire()
    b.add_not_gate(a, o)
    first = b.build()
    s, x, y, z = (b.add_wire() for _ in range(4))          # the builder is empty again, the switches are not
    b.add_multiplexer([x, y], s, z)
    second = b.build()
    assert first.info.cyclic == 0 and second.info.cyclic == 1   # a multiplexer that can float makes its net a resolved net


def test_pair_plan_statistics_need_no_device():
    """b2_sim_plan_pairs: the level-pair plan of a netlist, computed on the host; refused for netlists it cannot serve."""
    from digilogic_b200 import SimulatorBuilder, netgen
    from digilogic_b200.gsim import B200Error
    gen = netgen.random_dag(n_gates=20_000, depth=40, n_inputs=128, n_outputs=64, seed=7)
    sim = gen.emit(SimulatorBuilder()).build()
    plan = sim.plan_pairs(gen.outputs)
    assert plan["items"] + plan["children"] == 20_000                       # every gate once, as a parent or as a child
    assert 20 <= plan["launches"] <= 40 and plan["children"] > 6000 and plan["unstored"] > 1000
    assert plan["dropped_now"] + plan["dropped_later"] + plan["unstored"] < 20_000
    everything = sim.plan_pairs(np.arange(gen.n_wires, dtype=np.uint32))    # all wires kept: nothing may be dropped or skipped
    assert everything["unstored"] == 0 and everything["dropped_now"] == 0 and everything["dropped_later"] == 0
    assert everything["launches"] == plan["launches"] and everything["children"] == plan["children"]
    b = SimulatorBuilder()
    w = [b.add_wire() for _ in range(5)]
    b.add_and_gate([w[0], w[1], w[2]], w[3])                                # a three-input gate: not eligible
    b.add_not_gate(w[3], w[4])
    with pytest.raises(B200Error):
        b.build().plan_pairs([w[4]])
