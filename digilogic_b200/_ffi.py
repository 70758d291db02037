"""ctypes binding of include/digilogic_b200.h (the C ABI is the product boundary; this module
is the harness-side stub a maintainer's FFI layer would mirror — see INTEGRATION.md)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

u8, u32, u64, i32 = C.c_uint8, C.c_uint32, C.c_uint64, C.c_int
vp, sz = C.c_void_p, C.c_size_t
pu8, pu32, pu64 = C.POINTER(u8), C.POINTER(u32), C.POINTER(u64)


class SimInfo(C.Structure):
    _fields_ = [
        ("n_wires", u32), ("n_rows", u32), ("n_sources", u32), ("n_gates2", u32), ("n_gates_wide", u32),
        ("depth", u32), ("cyclic", u32), ("multi_driver", u32), ("n_lanes", u64), ("lane_words", u32),
        ("row_stride", u32), ("store_bytes", u64), ("last_mode", u32), ("n_gates3", u32),
        ("last_applications", u64),
    ]


class BatchInfo(C.Structure):
    _fields_ = [
        ("tile_words", u32), ("n_slots", u32), ("sweep", u32), ("tasks_per_pass", u32), ("grid_ctas", u32),
        ("gates_per_task", u32), ("n_ranks", u32), ("rank", u32), ("peers_mapped", u32), ("pair_launches", u32),
        ("store_bytes", u64), ("words_local", u64), ("last_tiles", u64), ("last_tickets", u64),
    ]


SIGNATURES = {
    "b2_last_error": (C.c_char_p, []),
    "b2_version": (C.c_char_p, []),
    "b2_kernel_launches": (u64, []),
    "b2_builder_new": (vp, []),
    "b2_builder_free": (None, [vp]),
    "b2_add_wire": (i32, [vp, u8, pu32]),
    "b2_add_wires": (i32, [vp, u8, u32, pu32]),
    "b2_builder_wire_width": (i32, [vp, u32, pu8]),
    "b2_add_gate": (i32, [vp, i32, pu32, sz, u32, pu32]),
    "b2_add_not": (i32, [vp, u32, u32, pu32]),
    "b2_add_mux": (i32, [vp, pu32, sz, u32, u32, pu32]),
    "b2_add_buffer": (i32, [vp, u32, u32, u32, pu32]),
    "b2_add_register": (i32, [vp, u32, u32, u32, u32, pu32]),
    "b2_add_horizontal_gate": (i32, [vp, i32, u32, u32, pu32]),
    "b2_add_add": (i32, [vp, u32, u32, u32, pu32]),
    "b2_add_sub": (i32, [vp, u32, u32, u32, pu32]),
    "b2_add_mul": (i32, [vp, u32, u32, u32, pu32]),
    "b2_add_compare": (i32, [vp, i32, u32, u32, u32, pu32]),
    "b2_add_shift": (i32, [vp, i32, u32, u32, u32, pu32]),
    "b2_add_slice": (i32, [vp, u32, u32, u32, pu32]),
    "b2_add_merge": (i32, [vp, pu32, sz, u32, pu32]),
    "b2_add_gates2": (i32, [vp, vp, vp, vp, vp, sz]),
    "b2_build": (vp, [vp]),
    "b2_sim_free": (None, [vp]),
    "b2_sim_get_info": (i32, [vp, C.POINTER(SimInfo)]),
    "b2_sim_get_pair_info": (i32, [vp, pu64, pu64, pu64]),
    "b2_sim_get_hot_rows": (i32, [vp, pu32, pu32]),
    "b2_sim_plan_pairs": (i32, [vp, vp, u32, pu64]),
    "b2_wire_width": (i32, [vp, u32, pu8]),
    "b2_sim_set_device": (i32, [vp, i32]),
    "b2_sim_set_stream": (i32, [vp, vp]),
    "b2_sim_synchronize": (i32, [vp]),
    "b2_sim_set_option": (i32, [vp, i32, C.c_int64]),
    "b2_sim_set_lanes": (i32, [vp, u64]),
    "b2_sim_max_lanes": (u64, [vp, u64]),
    "b2_set_wire_drive": (i32, [vp, u32, pu32, pu32, sz]),
    "b2_run_sim": (i32, [vp, u64]),
    "b2_get_wire_state": (i32, [vp, u32, pu32, pu32, sz]),
    "b2_set_wire_drive_lanes": (i32, [vp, u32, u32, u32, u32, vp, vp]),
    "b2_set_drives_lanes": (i32, [vp, vp, u32, vp, vp, sz]),
    "b2_fill_stimulus": (i32, [vp, vp, u32, u64, u64, i32]),
    "b2_run_sim_lanes": (i32, [vp, u64, vp, vp]),
    "b2_run_sim_lanes_async": (i32, [vp, u64]),
    "b2_get_wire_lanes": (i32, [vp, u32, u32, u32, u32, vp, vp]),
    "b2_gather_wires_device": (i32, [vp, vp, u32, vp, vp, sz]),
    "b2_gather_wires_host": (i32, [vp, vp, u32, vp, vp, sz]),
    "b2_gather_wires_host_async": (i32, [vp, vp, u32, vp, vp, sz]),
    "b2_pack_sim_state": (i32, [vp, vp, u32, vp, vp, sz, pu64]),
    "b2_builder_set_option": (i32, [vp, i32, C.c_int64]),
    "b2_sim_clone": (vp, [vp]),
    "b2_batch_new": (i32, [vp, vp, u32, vp, u32, u32, u32, C.POINTER(vp)]),
    "b2_batch_free": (None, [vp]),
    "b2_batch_set_device": (i32, [vp, i32]),
    "b2_batch_set_stream": (i32, [vp, vp]),
    "b2_batch_set_option": (i32, [vp, i32, C.c_int64]),
    "b2_batch_get_info": (i32, [vp, C.POINTER(BatchInfo)]),
    "b2_batch_run_generated": (i32, [vp, u64, i32, u64, u64, u64, vp, vp, sz, sz]),
    "b2_batch_run_planes": (i32, [vp, vp, vp, sz, sz, u64, u64, vp, vp, sz, sz]),
    "b2_batch_wait": (i32, [vp, C.POINTER(i32)]),
    "b2_comm_get_unique_id": (i32, [vp]),
    "b2_comm_init_rank": (i32, [C.POINTER(vp), i32, i32, vp, i32]),
    "b2_comm_free": (None, [vp]),
    "b2_batch_bind_comm": (i32, [vp, vp, u64]),
    "b2_batch_export_gather_handle": (i32, [vp, vp]),
    "b2_batch_import_gather_handles": (i32, [vp, vp]),
    "b2_batch_gather_outputs": (i32, [vp]),
    "b2_batch_gathered": (i32, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "b2_server_new": (vp, []),
    "b2_server_free": (None, [vp]),
    "b2_server_max_clients": (u64, [vp]),
    "b2_server_client_connected": (None, [vp, u64]),
    "b2_server_client_disconnected": (None, [vp, u64]),
    "b2_server_begin_build": (i32, [vp, u64]),
    "b2_server_end_build": (i32, [vp, u64]),
    "b2_server_add_net": (i32, [vp, u64, u8, pu32]),
    "b2_server_add_gate": (i32, [vp, u64, i32, u8, pu32, sz, u32, pu32]),
    "b2_server_add_not_gate": (i32, [vp, u64, u8, u32, u32, pu32]),
    "b2_server_add_mux": (i32, [vp, u64, u8, pu32, sz, u32, pu32]),
    "b2_server_set_net_drive": (i32, [vp, u64, u32, vp, sz, vp, sz]),
    "b2_server_eval": (i32, [vp, u64, u64]),
    "b2_server_get_net_state": (i32, [vp, u64, u32, pu8, C.POINTER(pu8), C.POINTER(pu8)]),
    "b2_server_sim_state": (i32, [vp, u64, vp, u32, vp, vp, sz, pu64]),
}

_lib = None


def lib():
    """Loads libdigilogic_b200.so, building it first if needed. Fails loudly: no fallback."""
    global _lib
    if _lib is None:
        path = os.environ.get("B2_LIB_PATH") or _build.LIB_PATH   # B2_LIB_PATH: A/B runs of two builds on one box
        if os.environ.get("B2_LIB_PATH"):
            pass
        elif not os.path.exists(path) or (os.environ.get("B2_AUTOBUILD", "1") == "1" and _build.needs_build()):
            path = _build.build()
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export the symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().b2_last_error().decode()
