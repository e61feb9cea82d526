"""ctypes front-end to the CHECKER libraries (test infrastructure only).

  * ``OracleLayer`` -> oracle/_build/libmcoracle.so  (oracle/mc_oracle.c, our C
    restatement of the reference path; always buildable, gcc only)
  * ``RefLayer``    -> oracle/_ref/libmcref.so       (the UNMODIFIED reference
    src/layer.cpp + src/random.cpp; buildable only where /root/reference
    exists, the built .so travels to the GPU box)

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
reference arm may import this module.  Nothing in mc_mpi_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libmcoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmcref.so")
REFERENCE_ROOT = os.environ.get("MCB200_REFERENCE_ROOT", "/root/reference")

# include/types/particle.hpp:7-18 -- the 24-byte wire format
PARTICLE_DTYPE = np.dtype(
    [("seed", "<u8"), ("x", "<f4"), ("mu", "<f4"), ("wmc", "<f4"), ("index", "<i4")]
)
assert PARTICLE_DTYPE.itemsize == 24


def _make(target: str) -> None:
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(
        ["make", "-s", "-C", HERE, target, f"REF={REFERENCE_ROOT}"],
        check=True, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
    )


def build_oracle() -> str:
    _make("oracle")
    return ORACLE_SO


def reference_sources_present() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "layer.cpp"))


def build_ref() -> str | None:
    """Compile the reference sources where they lie; None if they are absent
    and no prebuilt library travelled with the repo."""
    if reference_sources_present():
        _make("ref")
        try:   # the reference's own GPU prototype for sm_100a (a courtesy baseline, bench.py)
            _make("proto")
        except Exception:
            pass
    return REF_SO if os.path.isfile(REF_SO) else None


def have_ref() -> bool:
    return os.path.isfile(REF_SO) or reference_sources_present()


class _Stats(C.Structure):
    _fields_ = [
        ("events", C.c_int64), ("scatters", C.c_int64), ("n_left", C.c_int64),
        ("n_right", C.c_int64), ("n_dead", C.c_int64), ("w_left", C.c_double),
        ("w_right", C.c_double), ("w_dead", C.c_double),
    ]


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        P = C.c_void_p
        f, i = C.c_float, C.c_int
        sig = {
            "orc_rnd_real": (f, [C.POINTER(C.c_uint64)]),
            "orc_rnd_seed": (C.c_uint64, [C.POINTER(C.c_uint64)]),
            "orc_layer_new": (P, [f, f, i, i, f]),
            "orc_decompose_domain": (P, [f, f, f, i, i, i, i, f]),
            "orc_layer_free": (None, [P]),
            "orc_create_particles": (None, [P, f, f, i, C.c_uint64]),
            "orc_simulate": (None, [P, i]),
            "orc_simulate_mt": (None, [P, i, i]),
            "orc_nb_active": (i, [P]),
            "orc_dump_WA": (i, [P, C.c_char_p]),
            "orc_push": (None, [P, P, i]),
            "orc_tally_exact_lsb_log2": (i, []),
            "orc_tally_exact_f64": (None, [P, P]),
            "orc_class_weights_exact": (None, [P, P]),
            "orc_accumulate_exact": (C.c_double, [P, C.c_int64, P]),
            "orc_set_keep_border": (None, [P, i]),
            "orc_get_stats": (None, [P, C.POINTER(_Stats)]),
            "orc_clear_left": (None, [P]),
            "orc_clear_right": (None, [P]),
            "orc_logf_restated": (f, [f]),
            "orc_expf_restated": (f, [f]),
            "orc_logf_v": (None, [i, P, P, C.c_int64]),
            "orc_expf_v": (None, [i, P, P, C.c_int64]),
            "orc_rnd_real_v": (None, [C.POINTER(C.c_uint64), P, C.c_int64]),
            "orc_rnd_seed_v": (None, [C.POINTER(C.c_uint64), P, C.c_int64]),
        }
        for name in ("m", "index_start", "left_border", "right_border", "nb_disabled",
                     "nb_particles_create", "particles_size", "particles_left_size",
                     "particles_right_size", "absorbed_left_size", "absorbed_right_size",
                     "dead_size"):
            sig["orc_" + name] = (i, [P])
        for name in ("dx", "x_min", "x_max"):
            sig["orc_" + name] = (f, [P])
        for name in ("sigs", "absorption_rates", "weights_absorbed", "particles",
                     "particles_left", "particles_right", "tally_f64", "tally_exact",
                     "absorbed_left", "absorbed_right", "dead"):
            sig["orc_" + name] = (P, [P])
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _oracle = L
    return _oracle


def ref_lib():
    global _ref
    if _ref is None:
        so = build_ref()
        if so is None:
            raise RuntimeError("oracle/_ref/libmcref.so is absent and the reference sources "
                               f"are not at {REFERENCE_ROOT}")
        L = C.CDLL(so)
        P = C.c_void_p
        f, i = C.c_float, C.c_int
        sig = {
            "ref_decompose_domain": (P, [f, f, f, i, i, i, i, f]),
            "ref_layer_new": (P, [f, f, i, i, f]),
            "ref_layer_free": (None, [P]),
            "ref_create_particles": (None, [P, f, f, i, C.c_uint64]),
            "ref_simulate": (None, [P, i, i]),
            "ref_nb_active": (i, [P]),
            "ref_dump_WA": (None, [P]),
            "ref_push": (None, [P, P, i]),
            "ref_clear_left": (None, [P]),
            "ref_clear_right": (None, [P]),
            "ref_particle_step": (i, [P, P, P]),
            "ref_rnd_real": (f, [C.POINTER(C.c_uint64)]),
            "ref_rnd_seed": (C.c_uint64, [C.POINTER(C.c_uint64)]),
            "ref_sizeof_particle": (i, []),
        }
        for name in ("m", "left_border", "right_border", "nb_disabled", "nb_particles_create",
                     "particles_size", "particles_left_size", "particles_right_size"):
            sig["ref_" + name] = (i, [P])
        for name in ("dx", "x_min", "x_max"):
            sig["ref_" + name] = (f, [P])
        for name in ("sigs", "absorption_rates", "weights_absorbed", "particles",
                     "particles_left", "particles_right"):
            sig["ref_" + name] = (P, [P])
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _ref = L
    return _ref


def _view(ptr, n, dtype):
    """numpy view (no copy) of n items at a raw pointer; empty array for n==0."""
    dtype = np.dtype(dtype)
    if not ptr or n <= 0:
        return np.empty(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class _LayerBase:
    _prefix = ""
    _lib = None

    def _call(self, name, *a):
        return getattr(self._lib, self._prefix + name)(self._h, *a)

    def free(self):
        if self._h:
            self._call("layer_free")
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- public data of Layer (include/layer/layer.hpp:85-109) --
    m = property(lambda s: s._call("m"))
    dx = property(lambda s: s._call("dx"))
    x_min = property(lambda s: s._call("x_min"))
    x_max = property(lambda s: s._call("x_max"))
    left_border = property(lambda s: bool(s._call("left_border")))
    right_border = property(lambda s: bool(s._call("right_border")))
    nb_disabled = property(lambda s: s._call("nb_disabled"))
    nb_particles_create = property(lambda s: s._call("nb_particles_create"))
    sigs = property(lambda s: _view(s._call("sigs"), s.m, "<f4"))
    absorption_rates = property(lambda s: _view(s._call("absorption_rates"), s.m, "<f4"))
    weights_absorbed = property(lambda s: _view(s._call("weights_absorbed"), s.m, "<f4"))
    particles = property(lambda s: _view(s._call("particles"), s._call("particles_size"), PARTICLE_DTYPE))
    particles_left = property(
        lambda s: _view(s._call("particles_left"), s._call("particles_left_size"), PARTICLE_DTYPE))
    particles_right = property(
        lambda s: _view(s._call("particles_right"), s._call("particles_right_size"), PARTICLE_DTYPE))

    def nb_active(self):
        return self._call("nb_active")

    def create_particles(self, x_ini, wmc, n, seed):
        self._call("create_particles", x_ini, wmc, n, seed)

    def push(self, particles: np.ndarray):
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        if len(particles):
            self._call("push", particles.ctypes.data, len(particles))

    def clear_left(self):
        self._call("clear_left")

    def clear_right(self):
        self._call("clear_right")


class OracleLayer(_LayerBase):
    """oracle/mc_oracle.c through ctypes."""
    _prefix = "orc_"

    def __init__(self, handle):
        self._lib = oracle_lib()
        self._h = handle

    @classmethod
    def new(cls, x_min, x_max, index_start, m, particle_min_weight):
        return cls(oracle_lib().orc_layer_new(x_min, x_max, index_start, m, particle_min_weight))

    @classmethod
    def decompose_domain(cls, x_min, x_max, x_ini, world_size, world_rank, nb_cells,
                         nb_particles, particle_min_weight):
        return cls(oracle_lib().orc_decompose_domain(
            x_min, x_max, x_ini, world_size, world_rank, nb_cells, nb_particles,
            particle_min_weight))

    index_start = property(lambda s: s._call("index_start"))
    tally_f64 = property(lambda s: _view(s._call("tally_f64"), s.m, "<f8"))
    # exact tally as uint32[m, 4]: little-endian digits of the 128-bit sum per cell
    tally_exact = property(lambda s: _view(s._call("tally_exact"), 4 * s.m, "<u4").reshape(-1, 4))

    @property
    def tally_exact_f64(self):
        out = np.empty(self.m, dtype=np.float64)
        self._lib.orc_tally_exact_f64(self._h, out.ctypes.data)
        return out

    @property
    def class_weights_exact(self):
        """exact weight carried left / right / by the dead, rounded once to double"""
        out = np.empty(3, dtype=np.float64)
        self._lib.orc_class_weights_exact(self._h, out.ctypes.data)
        return out
    absorbed_left = property(
        lambda s: _view(s._call("absorbed_left"), s._call("absorbed_left_size"), PARTICLE_DTYPE))
    absorbed_right = property(
        lambda s: _view(s._call("absorbed_right"), s._call("absorbed_right_size"), PARTICLE_DTYPE))
    dead = property(lambda s: _view(s._call("dead"), s._call("dead_size"), PARTICLE_DTYPE))

    def set_keep_border(self, keep=True):
        self._call("set_keep_border", int(keep))

    def simulate(self, nb_particles, nthread=1):
        if nthread <= 1:
            self._call("simulate", nb_particles)
        else:
            self._call("simulate_mt", nb_particles, nthread)

    def dump_WA(self, path):
        if self._call("dump_WA", os.fsencode(path)) != 0:
            raise OSError(f"cannot write {path}")

    def stats(self) -> dict:
        st = _Stats()
        self._lib.orc_get_stats(self._h, C.byref(st))
        return {k: getattr(st, k) for k, _ in _Stats._fields_}


class RefLayer(_LayerBase):
    """The unmodified reference Layer (oracle/_ref/libmcref.so)."""
    _prefix = "ref_"

    def __init__(self, handle):
        self._lib = ref_lib()
        self._h = handle

    @classmethod
    def new(cls, x_min, x_max, index_start, m, particle_min_weight):
        return cls(ref_lib().ref_layer_new(x_min, x_max, index_start, m, particle_min_weight))

    @classmethod
    def decompose_domain(cls, x_min, x_max, x_ini, world_size, world_rank, nb_cells,
                         nb_particles, particle_min_weight):
        return cls(ref_lib().ref_decompose_domain(
            x_min, x_max, x_ini, world_size, world_rank, nb_cells, nb_particles,
            particle_min_weight))

    def simulate(self, nb_particles, nthread=1):
        self._call("simulate", nb_particles, nthread)

    def dump_WA(self):
        """writes ./WA.out in the current directory (src/layer.cpp:363-380)"""
        self._call("dump_WA")

    def particle_step(self, particle: np.ndarray, tally: np.ndarray) -> int:
        assert particle.dtype == PARTICLE_DTYPE and particle.shape == (1,)
        assert tally.dtype == np.float32 and len(tally) == self.m
        return self._call("particle_step", particle.ctypes.data, tally.ctypes.data)


def rnd_real(seed: int):
    s = C.c_uint64(seed)
    r = oracle_lib().orc_rnd_real(C.byref(s))
    return s.value, np.float32(r)


def rnd_seed(seed: int):
    s = C.c_uint64(seed)
    v = oracle_lib().orc_rnd_seed(C.byref(s))
    return s.value, v


def rnd_real_stream(seed: int, n: int):
    """n successive rnd_real draws; returns (final seed, float32 array)."""
    s = C.c_uint64(seed)
    out = np.empty(n, dtype=np.float32)
    oracle_lib().orc_rnd_real_v(C.byref(s), out.ctypes.data, n)
    return s.value, out


def rnd_seed_chain(seed: int, n: int):
    """first n values of the rnd_seed chain; returns (final state, uint64 array)."""
    s = C.c_uint64(seed)
    out = np.empty(n, dtype=np.uint64)
    oracle_lib().orc_rnd_seed_v(C.byref(s), out.ctypes.data, n)
    return s.value, out


def logf_v(x: np.ndarray, restated: bool) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    oracle_lib().orc_logf_v(int(restated), x.ctypes.data, out.ctypes.data, x.size)
    return out


def expf_v(x: np.ndarray, restated: bool) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    oracle_lib().orc_expf_v(int(restated), x.ctypes.data, out.ctypes.data, x.size)
    return out


def accumulate_exact(x: np.ndarray):
    """exact sum of floats: (uint32[4] little-endian digits, rounded double)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(4, dtype=np.uint32)
    d = oracle_lib().orc_accumulate_exact(x.ctypes.data, x.size, out.ctypes.data)
    return out, d
