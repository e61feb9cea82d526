"""Host-side mirror of the reference's Layer interface over the C ABI.

Same names and argument meaning as include/layer/layer.hpp / src/layer.cpp so
that the parity tests read like the reference's own tests.  Everything that
computes is in libmcb200.so (CUDA); this file only marshals.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _abi
from ._abi import PARTICLE_DTYPE, Counts, LayerDesc, check

f32 = np.float32
EPS_PRECISION = f32(1e-4)  # include/types/types.hpp:16
SEED0 = 5127801            # src/layer.cpp:36


class Layer:
    """One sub-slab on one GPU (Layer::Layer, src/layer.cpp:44-69)."""

    def __init__(self, x_min, x_max, index_start, m, particle_min_weight, *,
                 device=0, dx=None, sigs=None, absorption_rates=None, keep_border=False,
                 left_border=None, right_border=None):
        self._h = None
        self.x_min, self.x_max = f32(x_min), f32(x_max)
        self.index_start, self.m = int(index_start), int(m)
        self.particle_min_weight = f32(particle_min_weight)
        self.layer_dx = f32(self.x_max - self.x_min) / f32(self.m)  # Layer::dx, :47
        self.dx = f32(dx) if dx is not None else self.layer_dx      # edge dx actually tracked with
        self.device = int(device)
        d = LayerDesc()
        d.abi_version = _abi.ABI_VERSION
        d.device = self.device
        d.x_min, d.x_max = float(self.x_min), float(self.x_max)
        d.index_start, d.m = self.index_start, self.m
        d.dx = float(self.dx)
        d.particle_min_weight = float(self.particle_min_weight)
        d.left_border = -1 if left_border is None else int(bool(left_border))
        d.right_border = -1 if right_border is None else int(bool(right_border))
        keep = []
        for name, arr in (("sigs", sigs), ("absorption_rates", absorption_rates)):
            if arr is None:
                setattr(d, name, None)
            else:
                a = np.ascontiguousarray(arr, dtype=np.float32)
                if a.shape != (self.m,):
                    raise ValueError(f"{name} must have m={self.m} entries")
                keep.append(a)
                setattr(d, name, a.ctypes.data)
        d.keep_border = int(bool(keep_border))
        h = C.c_void_p()
        check(_abi.lib().mcb200_layer_create(C.byref(d), C.byref(h)))
        self._h = h
        self.left_border = (abs(float(self.x_min)) < float(EPS_PRECISION)
                            if left_border is None else bool(left_border))
        self.right_border = (abs(float(self.x_max) - 1.0) < float(EPS_PRECISION)
                             if right_border is None else bool(right_border))

    # -- life cycle --
    def clone(self) -> "Layer":
        """deep copy, device state included (Layer is copied / returned by value in the
        reference: src/layer.cpp:41, include/mcmpi/worker.hpp:60)"""
        h = C.c_void_p()
        check(_abi.lib().mcb200_layer_clone(self._h, C.byref(h)))
        other = object.__new__(Layer)
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_h"})
        other._h = h
        return other

    def close(self):
        if self._h is not None:
            _abi.lib().mcb200_layer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- sources --
    def create_particles(self, x_ini, wmc, n, seed=SEED0):
        """Layer::create_particles, src/layer.cpp:71-82 (births happen on the device)."""
        check(_abi.lib().mcb200_layer_create_particles(self._h, float(x_ini), float(wmc), int(n),
                                                       int(seed)))

    def push(self, particles: np.ndarray):
        """append host particles to the bank (the workers' receive)."""
        p = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        if len(p):
            check(_abi.lib().mcb200_layer_push(self._h, p.ctypes.data, len(p)))

    def push_device(self, dev_ptr: int, n: int):
        check(_abi.lib().mcb200_layer_push_device(self._h, C.c_void_p(dev_ptr), int(n)))

    def simulate_host(self, particles) -> dict:
        """track particles that sit in host memory (a numpy array of PARTICLE_DTYPE, or a
        (pointer, count) pair for pinned buffers): chunked, the H2D copy of chunk k+1 under the
        tracking of chunk k -- what Layer::simulate(.., use_gpu) does with `particles`"""
        c = Counts()
        if isinstance(particles, tuple):
            ptr, n = particles
        else:
            p = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
            ptr, n = p.ctypes.data, len(p)
        check(_abi.lib().mcb200_layer_simulate_host(self._h, C.c_void_p(ptr), int(n), C.byref(c)))
        return c.as_dict()

    def reset_tally(self):
        check(_abi.lib().mcb200_layer_reset_tally(self._h))

    # -- the hot path --
    def simulate(self, nb_particles=-1) -> dict:
        """Layer::simulate, src/layer.cpp:239-361; -1 = until nothing is left."""
        c = Counts()
        check(_abi.lib().mcb200_layer_simulate(self._h, int(nb_particles), C.byref(c)))
        return c.as_dict()

    # -- results --
    def counts(self) -> dict:
        c = Counts()
        check(_abi.lib().mcb200_layer_counts(self._h, C.byref(c)))
        return c.as_dict()

    nb_disabled = property(lambda s: s.counts()["nb_disabled"])

    def nb_active(self) -> int:
        return self.counts()["nb_active"]

    def _pop(self, fn, n):
        out = np.empty(max(int(n), 0), dtype=PARTICLE_DTYPE)
        got = C.c_int64(0)
        check(fn(self._h, out.ctypes.data if len(out) else None, len(out), C.byref(got)))
        return out[: got.value]

    def pop_left(self) -> np.ndarray:
        """particles_left (layer.hpp:94), cleared like the workers do after sending."""
        return self._pop(_abi.lib().mcb200_layer_pop_left, self.counts()["n_outbox_left"])

    def pop_right(self) -> np.ndarray:
        return self._pop(_abi.lib().mcb200_layer_pop_right, self.counts()["n_outbox_right"])

    def pop_device(self, side: int, dev_ptr: int, cap: int) -> int:
        fn = (_abi.lib().mcb200_layer_pop_left_device if side == 0
              else _abi.lib().mcb200_layer_pop_right_device)
        got = C.c_int64(0)
        check(fn(self._h, C.c_void_p(dev_ptr), int(cap), C.byref(got)))
        return got.value

    def outbox_device(self, side: int):
        """(device pointer, n): the outbox itself, n contiguous 24-byte records (zero-copy;
        valid until the next simulate())."""
        ptr, n = C.c_void_p(), C.c_int64(0)
        check(_abi.lib().mcb200_layer_outbox_device(self._h, int(side), C.byref(ptr), C.byref(n)))
        return int(ptr.value or 0), n.value

    def outbox_clear(self, side: int):
        check(_abi.lib().mcb200_layer_outbox_clear(self._h, int(side)))

    # -- direct peer exchange over NVLink (mcb200.h) --
    def inbox_create(self, max_take: int):
        """allocate this layer's inbox; -> (IPC handle bytes, geometry bytes) for the peers"""
        handle = (C.c_uint8 * _abi.IPC_HANDLE_BYTES)()
        geom = _abi.InboxGeom()
        check(_abi.lib().mcb200_layer_inbox_create(self._h, int(max_take), handle, C.byref(geom)))
        return bytes(handle), bytes(geom)

    def connect_peer(self, side: int, handle: bytes, geom: bytes):
        """route escapees on `side` into the inbox of a neighbour living in another process"""
        h = (C.c_uint8 * _abi.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        g = _abi.InboxGeom.from_buffer_copy(geom)
        check(_abi.lib().mcb200_layer_connect_peer(self._h, int(side), h, C.byref(g)))

    def connect_local(self, side: int, other: "Layer"):
        """same, the neighbour being a layer of this process"""
        check(_abi.lib().mcb200_layer_connect_local(self._h, int(side), other._h))

    def disconnect_peers(self):
        check(_abi.lib().mcb200_layer_disconnect_peers(self._h))

    def set_exchange_parity(self, parity: int):
        check(_abi.lib().mcb200_layer_set_exchange_parity(self._h, int(parity)))

    def ingest_inbox(self, from_side: int, parity: int) -> int:
        n = C.c_int64(0)
        check(_abi.lib().mcb200_layer_ingest_inbox(self._h, int(from_side), int(parity), C.byref(n)))
        return n.value

    @property
    def weights_absorbed(self) -> np.ndarray:
        out = np.empty(self.m, dtype=np.float32)
        check(_abi.lib().mcb200_layer_weights_absorbed(self._h, out.ctypes.data))
        return out

    @property
    def weights_absorbed_f64(self) -> np.ndarray:
        out = np.empty(self.m, dtype=np.float64)
        check(_abi.lib().mcb200_layer_weights_absorbed_f64(self._h, out.ctypes.data))
        return out

    def weights_absorbed_exact(self):
        """(uint32[m, 4], lsb_log2): the exact tally, little-endian digits of a 128-bit
        two's-complement integer per cell; tally == integer * 2**lsb_log2."""
        out = np.empty((self.m, 4), dtype=np.uint32)
        k = C.c_int32(0)
        check(_abi.lib().mcb200_layer_weights_absorbed_exact(self._h, out.ctypes.data,
                                                             C.byref(k)))
        return out, k.value

    def dump_WA(self, path="WA.out"):
        """Layer::dump_WA, src/layer.cpp:363-380."""
        check(_abi.lib().mcb200_layer_dump_WA(self._h, str(path).encode()))

    # -- cross-sections (public mutable vectors in the reference, layer.hpp:103-104) --
    def set_cross_sections(self, sigs=None, absorption_rates=None):
        s = None if sigs is None else np.ascontiguousarray(sigs, dtype=np.float32)
        a = None if absorption_rates is None else np.ascontiguousarray(absorption_rates, np.float32)
        for arr in (s, a):
            if arr is not None and arr.shape != (self.m,):
                raise ValueError(f"cross-section tables must have m={self.m} entries")
        check(_abi.lib().mcb200_layer_set_cross_sections(
            self._h, None if s is None else s.ctypes.data, None if a is None else a.ctypes.data))

    def get_cross_sections(self):
        s = np.empty(self.m, dtype=np.float32)
        a = np.empty(self.m, dtype=np.float32)
        check(_abi.lib().mcb200_layer_get_cross_sections(self._h, s.ctypes.data, a.ctypes.data))
        return s, a

    sigs = property(lambda self: self.get_cross_sections()[0])
    absorption_rates = property(lambda self: self.get_cross_sections()[1])

    # -- plumbing --
    def set_option(self, key: str, value: int):
        check(_abi.lib().mcb200_layer_set_option(self._h, key.encode(), int(value)))

    @property
    def stream_ptr(self) -> int:
        return int(_abi.lib().mcb200_layer_stream(self._h) or 0)


def default_cross_sections(x_min, x_max, m):
    """(sigs, absorption_rates) Layer::Layer hard-codes for m cells on [x_min, x_max]
    (src/layer.cpp:53-63), computed by the library with the libm the reference links."""
    s = np.empty(int(m), dtype=np.float32)
    a = np.empty(int(m), dtype=np.float32)
    check(_abi.lib().mcb200_default_cross_sections(float(f32(x_min)), float(f32(x_max)), int(m),
                                                   s.ctypes.data, a.ctypes.data))
    return s, a


def split_cells(nb_cells: int, world_size: int, world_rank: int):
    """cells of one rank, src/layer.cpp:24-27 -> (start_index, nb_my_cells)."""
    cells_per_layer = nb_cells // world_size
    num_with_extra = nb_cells % world_size
    nb_my_cells = cells_per_layer + (1 if world_rank < num_with_extra else 0)
    start_index = world_rank * cells_per_layer + min(world_rank, num_with_extra)
    return start_index, nb_my_cells


def decompose_domain(x_min, x_max, x_ini, world_size, world_rank, nb_cells, nb_particles,
                     particle_min_weight, *, device=0, global_dx=False, keep_border=False,
                     sigs=None, absorption_rates=None, seed=SEED0, cells=None,
                     left_border=None, right_border=None) -> Layer:
    """decompose_domain, src/layer.cpp:17-42, float arithmetic mirrored in float32.

    global_dx=False reproduces the reference (every layer recomputes its own dx
    from its rounded bounds, :47); global_dx=True tracks every sub-slab with the
    one global dx so that K GPUs give the single-GPU trajectories bit for bit.
    `sigs` / `absorption_rates`, if given, are GLOBAL tables (nb_cells entries).
    `cells` = (start_index, nb_my_cells) overrides the reference's equal split
    (:24-27) -- with global_dx the result does not depend on where the cuts are,
    so a driver may place them where the work balances.
    """
    x_min, x_max, x_ini = f32(x_min), f32(x_max), f32(x_ini)
    start_index, nb_my_cells = (split_cells(nb_cells, world_size, world_rank) if cells is None
                                else (int(cells[0]), int(cells[1])))
    dx = f32(x_max - x_min) / f32(nb_cells)                      # :29
    cell_ini = int(f32(x_ini - x_min) / dx)                      # :30
    lo = f32(x_min + f32(start_index) * dx)                      # :32
    hi = f32(x_min + f32(start_index + nb_my_cells) * dx)
    wmc = f32(1.0 / nb_particles)                                # :38, double -> float
    sl = slice(start_index, start_index + nb_my_cells)
    if global_dx and (sigs is None or absorption_rates is None):
        # the table of the WHOLE slab, sliced: per-sub-slab recomputation (the reference's
        # K-rank behaviour) moves a quarter of the entries by 1 ulp
        s_all, a_all = default_cross_sections(x_min, x_max, nb_cells)
        sigs = s_all if sigs is None else sigs
        absorption_rates = a_all if absorption_rates is None else absorption_rates
    layer = Layer(lo, hi, start_index, nb_my_cells, particle_min_weight,
                  device=device, dx=dx if global_dx else None, keep_border=keep_border,
                  left_border=left_border, right_border=right_border,
                  sigs=None if sigs is None else np.asarray(sigs, dtype=np.float32)[sl],
                  absorption_rates=(None if absorption_rates is None
                                    else np.asarray(absorption_rates, dtype=np.float32)[sl]))
    if start_index <= cell_ini < start_index + nb_my_cells:      # :34
        layer.create_particles(x_ini, wmc, nb_particles, seed)
    return layer
