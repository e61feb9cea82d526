"""Device primitives vs the CPU oracle / libm, through the C ABI (needs a B200)."""
import ctypes as C

import numpy as np
import pytest

from mc_mpi_b200 import _abi, configs
from oracle import pyoracle
from oracle.pyoracle import PARTICLE_DTYPE, OracleLayer

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_rnd_real_device_equals_cpu(gpu, mcb_lib):
    """The reference's TestCudaRandom (src/test_curandom.cu:12-82): seeds from the rnd_seed
    chain starting 30061994, CPU rnd_real vs the device kernel, EXACT equality of the
    advanced seeds and of the floats."""
    n = 2_000_000
    _, seeds = pyoracle.rnd_seed_chain(30061994, n)
    dev_seeds = seeds.copy()
    dev_out = np.empty(n, dtype=np.float32)
    _abi.check(mcb_lib.mcb200_test_rnd_real(gpu, dev_seeds.ctypes.data, dev_out.ctypes.data, n))
    g, c, mask = np.uint64(6364136223846793005), np.uint64(1442695040888963407), np.uint64((1 << 63) - 1)
    with np.errstate(over="ignore"):
        cpu_seeds = (g * seeds + c) & mask
    assert np.array_equal(dev_seeds, cpu_seeds)
    # the float: through the oracle's C rnd_real on a sample (python loop), numpy for all
    cpu_out = cpu_seeds.astype(np.float32) * np.float32(2.0 ** -63)
    assert np.array_equal(bits(dev_out), bits(cpu_out))
    for i in range(0, n, n // 50):
        s, r = pyoracle.rnd_real(int(seeds[i]))
        assert s == int(dev_seeds[i]) and r.view(np.uint32) == dev_out[i].view(np.uint32)
    # edge states: 0, 1, 2^63-1 and the state whose successor rounds up to 1.0f
    edge = np.array([0, 1, (1 << 63) - 1, 5127801], dtype=np.uint64)
    e_dev = edge.copy()
    e_out = np.empty(len(edge), dtype=np.float32)
    _abi.check(mcb_lib.mcb200_test_rnd_real(gpu, e_dev.ctypes.data, e_out.ctypes.data, len(edge)))
    for i, s0 in enumerate(edge.tolist()):
        s, r = pyoracle.rnd_real(s0)
        assert s == int(e_dev[i]) and r.view(np.uint32) == e_out[i].view(np.uint32)


def test_device_logf_equals_libm(gpu, mcb_lib):
    """device logf (FP64 restatement of glibc's algorithm) == the libm the reference links,
    bit for bit, over the path's domain h in {0} U [2^-63, 1]."""
    rng = np.random.default_rng(7)
    h = np.concatenate([
        rng.integers(0, 1 << 63, size=3_000_000, dtype=np.uint64).astype(np.float32) * np.float32(2.0 ** -63),
        rng.integers(0x1f800000, 0x3f800001, size=3_000_000, dtype=np.uint32).view(np.float32),
        # every float in [0.5, 1): where cancellation would show
        np.arange(0x3f000000, 0x3f800001, 3, dtype=np.uint32).view(np.float32),
        np.array([0.0, 1.0, 2.0 ** -63, 0.5], dtype=np.float32),
    ]).astype(np.float32)
    out = np.empty_like(h)
    _abi.check(mcb_lib.mcb200_test_logf(gpu, h.ctypes.data, out.ctypes.data, h.size))
    want = pyoracle.logf_v(h, restated=False)
    assert np.array_equal(bits(out), bits(want))


def test_device_expf_equals_libm(gpu, mcb_lib):
    """device expf on [-inf, +0]; the path consumes 1 - expf(), which must match exactly."""
    rng = np.random.default_rng(8)
    x = np.concatenate([
        -rng.integers(0, 0x7f800001, size=3_000_000, dtype=np.uint32).view(np.float32),
        -rng.random(3_000_000, dtype=np.float32) * np.float32(30),
        -rng.random(1_000_000, dtype=np.float32) * np.float32(1e-3),
        np.array([0.0, -0.0, -np.inf, -88.0, -103.9, -104.5, -1e-30, -1e-40, -3.4e38], dtype=np.float32),
    ]).astype(np.float32)
    out = np.empty_like(x)
    _abi.check(mcb_lib.mcb200_test_expf(gpu, x.ctypes.data, out.ctypes.data, x.size))
    want = pyoracle.expf_v(x, restated=False)
    one = np.float32(1)
    assert np.array_equal(bits(one - out), bits(one - want))
    # the raw values agree too, except possibly in the flush-to-zero tail below 2^-126
    big = want >= np.float32(1.2e-38)
    assert np.count_nonzero(bits(out[big]) != bits(want[big])) <= 1


@pytest.mark.parametrize("n", [1, 31, 1000, 300_001])
def test_device_birth_equals_reference_create_particles(gpu, mcb_lib, n):
    """Layer::create_particles (src/layer.cpp:89-121): seed chain by LCG jump-ahead, first draw
    -> mu.  Compared with the oracle's sequential births, bit for bit, in chain order."""
    cfg = configs.reference_default(n)
    o = OracleLayer.decompose_domain(cfg.x_min, cfg.x_max, cfg.x_ini, 1, 0, cfg.nb_cells, n,
                                     cfg.particle_min_weight)
    want = o.particles.copy()
    assert len(want) == n
    out = np.empty(n, dtype=PARTICLE_DTYPE)
    wmc = np.float32(1.0 / n)
    _abi.check(mcb_lib.mcb200_test_birth(gpu, cfg.x_ini, float(wmc), float(o.dx), n, 5127801,
                                         out.ctypes.data))
    assert out.tobytes() == want.tobytes()


def test_exact_accumulator_device_equals_checker(gpu, mcb_lib):
    """the long accumulator the tally uses: device sum of floats == the checker's 128-bit
    integer sum, digit for digit, including negative deposits, carries and the dropped
    sub-LSB bits."""
    rng = np.random.default_rng(11)
    for x in (
        rng.random(1_000_000, dtype=np.float32) * np.float32(1e-5),
        rng.integers(0x0d800000, 0x3f000000, size=1_000_000, dtype=np.uint32).view(np.float32),
        np.concatenate([rng.random(100_000, dtype=np.float32), -rng.random(100_000, dtype=np.float32)]),
        np.array([0.0, -0.0, 1e-30, 2.0 ** -97, 2.0 ** -110, 2.0 ** -125, 1e-45, 100.0, -100.0, 3e-39],
                 dtype=np.float32),
        np.full(300_000, 127.99 / 300_000 * 100, dtype=np.float32)[:3000],
    ):
        x = np.ascontiguousarray(x, dtype=np.float32)
        want, want_d = pyoracle.accumulate_exact(x)
        got = np.zeros(4, dtype=np.uint32)
        d = C.c_double(0)
        _abi.check(mcb_lib.mcb200_test_accumulate(gpu, x.ctypes.data, x.size, got.ctypes.data,
                                                  C.byref(d)))
        assert np.array_equal(got, want)
        assert d.value == want_d
    bad = np.array([1.0, 200.0], dtype=np.float32)
    got = np.zeros(4, dtype=np.uint32)
    assert mcb_lib.mcb200_test_accumulate(gpu, bad.ctypes.data, 2, got.ctypes.data, None) == _abi.ERR_RANGE


def test_edge_distance_division_is_ieee(gpu, mcb_lib):
    """di_edge = (x_edge - x)/mu with the reciprocal of mu hoisted out of the event loop must be
    the IEEE round-to-nearest quotient the reference's x86 divss gives, for every operand."""
    rng = np.random.default_rng(21)
    n = 4_000_000
    mu = (rng.random(n, dtype=np.float32) * np.float32(2) - np.float32(1)).astype(np.float32)
    a = np.concatenate([
        (rng.random(n // 2, dtype=np.float32) * np.float32(0.01)).astype(np.float32),      # cell-sized
        rng.integers(0, 0x7f800000, size=n // 4, dtype=np.uint32).view(np.float32),          # any exponent
        -rng.integers(0, 0x7f800000, size=n // 4, dtype=np.uint32).view(np.float32),
    ]).astype(np.float32)
    # special operands: zeros, denormals, huge / tiny mu, +-EPS
    sa = np.array([0.0, -0.0, 1e-45, 1e-38, 1.0, 3.4e38, 1e-30, 0.001, 0.001, 0.001, 0.001, 0.5],
                  dtype=np.float32)
    sm = np.array([0.5, 0.5, 0.5, 0.5, 1e-4, 0.5, 1.0, 1e-4, 1.0001e-4, -1.0001e-4, 0.0, 3e30],
                  dtype=np.float32)
    a, mu = np.concatenate([a, sa]), np.concatenate([mu, sm])
    out = np.empty_like(a)
    _abi.check(mcb_lib.mcb200_test_edge_distance(gpu, a.ctypes.data, mu.ctypes.data,
                                                 out.ctypes.data, a.size))
    eps = np.float32(1e-4)
    with np.errstate(all="ignore"):
        want = np.where((mu < -eps) | (eps < mu), a / mu, np.float32(3.402823466e+38)).astype(np.float32)
    assert np.array_equal(bits(out), bits(want))
