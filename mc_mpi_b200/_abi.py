"""ctypes binding of include/mcb200.h (libmcb200.so).

The product path has no CPU fallback: if the CUDA library is missing or was
not built, importing the binding raises -- loudly -- instead of degrading.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCB200_LIB") or os.path.join(HERE, "libmcb200.so")

ABI_VERSION = 1
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY, ERR_RANGE, ERR_TIMEOUT = 0, -1, -2, -3, -4, -5, -6

# include/types/particle.hpp:7-18 -- 24-byte wire format
PARTICLE_DTYPE = np.dtype(
    [("seed", "<u8"), ("x", "<f4"), ("mu", "<f4"), ("wmc", "<f4"), ("index", "<i4")]
)
assert PARTICLE_DTYPE.itemsize == 24


class McbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mcb200 error {code}: {msg}")
        self.code = code


class LayerDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("x_min", C.c_float), ("x_max", C.c_float),
        ("index_start", C.c_int32), ("m", C.c_int32),
        ("dx", C.c_float), ("particle_min_weight", C.c_float),
        ("left_border", C.c_int32), ("right_border", C.c_int32),
        ("sigs", C.c_void_p), ("absorption_rates", C.c_void_p),
        ("keep_border", C.c_int32),
    ]


class Counts(C.Structure):
    _fields_ = [
        ("nb_disabled", C.c_int64), ("nb_active", C.c_int64), ("n_bank", C.c_int64),
        ("n_unborn", C.c_int64), ("n_outbox_left", C.c_int64), ("n_outbox_right", C.c_int64),
        ("events", C.c_int64), ("scatters", C.c_int64), ("n_left", C.c_int64),
        ("n_right", C.c_int64), ("n_dead", C.c_int64), ("w_left", C.c_double),
        ("w_right", C.c_double), ("w_dead", C.c_double), ("launches", C.c_int64),
        ("track_ms", C.c_double), ("gpu_launches", C.c_int64),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


IPC_HANDLE_BYTES = 64


class InboxGeom(C.Structure):
    _fields_ = [
        ("nstripes", C.c_int32), ("stripe_cap", C.c_int32), ("ovf_cap", C.c_int64),
        ("max_take", C.c_int64), ("slot_bytes", C.c_int64), ("fills_offset", C.c_int64),
    ]


class WorldDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("world_size", C.c_int32),
        ("x_min", C.c_float), ("x_max", C.c_float), ("x_ini", C.c_float),
        ("nb_cells", C.c_int32), ("particle_min_weight", C.c_float),
        ("cuts", C.c_void_p), ("sigs", C.c_void_p), ("absorption_rates", C.c_void_p),
        ("windows", C.c_int32), ("block", C.c_int32), ("max_ctas", C.c_int32),
        ("ring_cap", C.c_int32), ("retire_batch", C.c_int32), ("xs_global", C.c_int32),
        ("bank_cap", C.c_int64), ("inflight_limit", C.c_int64),
    ]


class WorldGeom(C.Structure):
    _fields_ = [
        ("rank", C.c_int32), ("world_size", C.c_int32), ("stripes", C.c_int32),
        ("ring_cap", C.c_int32), ("block_bytes", C.c_int64),
        ("off_rec", C.c_int64 * 2), ("off_credit", C.c_int64 * 2), ("off_chain", C.c_int64),
    ]


class WorldResult(C.Structure):
    _fields_ = [
        ("events", C.c_int64), ("scatters", C.c_int64),
        ("n_left", C.c_int64), ("n_right", C.c_int64), ("n_dead", C.c_int64),
        ("births", C.c_int64), ("sent_left", C.c_int64), ("sent_right", C.c_int64),
        ("window_crossings", C.c_int64),
        ("idle_polls", C.c_int64), ("blocked_passes", C.c_int64),
        ("bank_pushes", C.c_int64), ("bank_pops", C.c_int64), ("lane_slots", C.c_int64),
        ("idle_warp_ns", C.c_int64),
        ("w_left", C.c_double), ("w_right", C.c_double), ("w_dead", C.c_double),
        ("kernel_ms", C.c_double),
        ("windows", C.c_int32), ("ctas", C.c_int32), ("block", C.c_int32),
        ("stripes", C.c_int32), ("ring_cap", C.c_int32), ("error", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/mcb200.h declares: name -> (restype, argtypes)
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "mcb200_layer_create": (C.c_int, [C.POINTER(LayerDesc), C.POINTER(_P)]),
    "mcb200_layer_destroy": (None, [_P]),
    "mcb200_layer_clone": (C.c_int, [_P, C.POINTER(_P)]),
    "mcb200_layer_set_cross_sections": (C.c_int, [_P, _P, _P]),
    "mcb200_layer_get_cross_sections": (C.c_int, [_P, _P, _P]),
    "mcb200_default_cross_sections": (C.c_int, [_F, _F, _I32, _P, _P]),
    "mcb200_layer_create_particles": (C.c_int, [_P, _F, _F, _I64, C.c_uint64]),
    "mcb200_layer_push": (C.c_int, [_P, _P, _I64]),
    "mcb200_layer_push_device": (C.c_int, [_P, _P, _I64]),
    "mcb200_layer_simulate": (C.c_int, [_P, _I64, C.POINTER(Counts)]),
    "mcb200_layer_simulate_host": (C.c_int, [_P, _P, _I64, C.POINTER(Counts)]),
    "mcb200_layer_reset_tally": (C.c_int, [_P]),
    "mcb200_layer_counts": (C.c_int, [_P, C.POINTER(Counts)]),
    "mcb200_layer_pop_left": (C.c_int, [_P, _P, _I64, C.POINTER(_I64)]),
    "mcb200_layer_pop_right": (C.c_int, [_P, _P, _I64, C.POINTER(_I64)]),
    "mcb200_layer_pop_left_device": (C.c_int, [_P, _P, _I64, C.POINTER(_I64)]),
    "mcb200_layer_pop_right_device": (C.c_int, [_P, _P, _I64, C.POINTER(_I64)]),
    "mcb200_layer_outbox_device": (C.c_int, [_P, _I32, C.POINTER(_P), C.POINTER(_I64)]),
    "mcb200_layer_outbox_clear": (C.c_int, [_P, _I32]),
    "mcb200_layer_inbox_create": (C.c_int, [_P, _I64, _P, _P]),
    "mcb200_layer_connect_peer": (C.c_int, [_P, _I32, _P, _P]),
    "mcb200_layer_connect_local": (C.c_int, [_P, _I32, _P]),
    "mcb200_layer_disconnect_peers": (C.c_int, [_P]),
    "mcb200_layer_set_exchange_parity": (C.c_int, [_P, _I32]),
    "mcb200_layer_ingest_inbox": (C.c_int, [_P, _I32, _I32, C.POINTER(_I64)]),
    "mcb200_world_create": (C.c_int, [C.POINTER(WorldDesc), C.POINTER(_P)]),
    "mcb200_world_destroy": (None, [_P]),
    "mcb200_world_export": (C.c_int, [_P, _P, C.POINTER(WorldGeom)]),
    "mcb200_world_connect_peer": (C.c_int, [_P, _I32, _P, C.POINTER(WorldGeom)]),
    "mcb200_world_connect_local": (C.c_int, [_P, _P]),
    "mcb200_world_disconnect": (C.c_int, [_P]),
    "mcb200_world_prepare": (C.c_int, [_P, _I64, C.c_uint64]),
    "mcb200_world_launch": (C.c_int, [_P]),
    "mcb200_world_wait": (C.c_int, [_P, C.POINTER(WorldResult)]),
    "mcb200_world_run": (C.c_int, [C.POINTER(_P), _I32, _I64, C.c_uint64, C.POINTER(WorldResult)]),
    "mcb200_world_cells": (C.c_int, [_P, C.POINTER(_I32), C.POINTER(_I32)]),
    "mcb200_world_tally": (C.c_int, [_P, _P]),
    "mcb200_world_tally_f64": (C.c_int, [_P, _P]),
    "mcb200_world_tally_exact": (C.c_int, [_P, _P, C.POINTER(_I32)]),
    "mcb200_world_reset_tally": (C.c_int, [_P]),
    "mcb200_world_gather_tally_f64": (C.c_int, [C.POINTER(_P), _I32, _P]),
    "mcb200_world_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
    "mcb200_world_stream": (_P, [_P]),
    "mcb200_layer_weights_absorbed": (C.c_int, [_P, _P]),
    "mcb200_layer_weights_absorbed_f64": (C.c_int, [_P, _P]),
    "mcb200_layer_weights_absorbed_exact": (C.c_int, [_P, _P, C.POINTER(_I32)]),
    "mcb200_layer_dump_WA": (C.c_int, [_P, C.c_char_p]),
    "mcb200_layer_stream": (_P, [_P]),
    "mcb200_layer_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
    "mcb200_last_error": (C.c_char_p, []),
    "mcb200_abi_version": (C.c_int, []),
    "mcb200_device_count": (C.c_int, []),
    "mcb200_test_rnd_real": (C.c_int, [C.c_int, _P, _P, _I64]),
    "mcb200_test_philox": (C.c_int, [C.c_int, _P, _P, _P, _P, _I64]),
    "mcb200_test_logf": (C.c_int, [C.c_int, _P, _P, _I64]),
    "mcb200_test_expf": (C.c_int, [C.c_int, _P, _P, _I64]),
    "mcb200_test_edge_distance": (C.c_int, [C.c_int, _P, _P, _P, _I64]),
    "mcb200_test_accumulate": (C.c_int, [C.c_int, _P, _I64, _P, C.POINTER(C.c_double)]),
    "mcb200_test_birth": (C.c_int, [C.c_int, _F, _F, _F, _I64, C.c_uint64, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded libmcb200.so; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m mc_mpi_b200.build` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError = header / library mismatch
            fn.restype, fn.argtypes = res, args
        if L.mcb200_abi_version() != ABI_VERSION:
            raise ImportError("libmcb200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise McbError(rc, lib().mcb200_last_error().decode(errors="replace"))


def device_count() -> int:
    return lib().mcb200_device_count()
