"""The C-ABI library loads on a CPU-only box and exports every symbol include/mcb200.h
declares; without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcb200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb200_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(mcb_lib):
    from mc_mpi_b200 import _abi
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(mcb_lib, n), f"{n} declared in mcb200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_abi.SYMBOLS) == names
    assert mcb_lib.mcb200_abi_version() == _abi.ABI_VERSION


def test_reference_operator_symbols_exported(mcb_lib):
    # the reference's existing operator boundary, include/culayer/culayer.hpp:6-13, and
    # decompose_domain (include/layer/layer.hpp:112-114), with the reference's C++ mangling
    from mc_mpi_b200 import _abi
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], check=True,
                         stdout=subprocess.PIPE, text=True).stdout
    assert "_Z10cusimulateiP12particle_tagPKfS2_Pfiif" in out
    assert "_Z16decompose_domainfffiiiif" in out
    assert "_ZN5Layer8simulateEiib" in out
    assert "_ZN5Layer16create_particlesEffiy" in out
    assert "_ZN5Layer7dump_WAEv" in out
    assert "_ZNK5Layer9nb_activeEv" in out


def test_struct_layouts_match_header(mcb_lib):
    from mc_mpi_b200 import _abi
    # compile a tiny C program against the header and compare sizeof/offsetof
    code = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "mcb200.h"
    int main(void){
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(mcb200_particle), sizeof(mcb200_layer_desc),
             sizeof(mcb200_counts), offsetof(mcb200_layer_desc, sigs),
             offsetof(mcb200_layer_desc, keep_border), offsetof(mcb200_counts, track_ms));
      return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(code)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        src, "-o", exe], check=True)
        got = list(map(int, subprocess.run([exe], check=True, stdout=subprocess.PIPE,
                                           text=True).stdout.split()))
    want = [24, C.sizeof(_abi.LayerDesc), C.sizeof(_abi.Counts), _abi.LayerDesc.sigs.offset,
            _abi.LayerDesc.keep_border.offset, _abi.Counts.track_ms.offset]
    assert got == want


def test_argument_validation_without_compute(mcb_lib):
    from mc_mpi_b200 import _abi
    h = C.c_void_p()
    d = _abi.LayerDesc()
    d.abi_version = 999
    assert mcb_lib.mcb200_layer_create(C.byref(d), C.byref(h)) == _abi.ERR_INVALID
    assert b"abi_version" in mcb_lib.mcb200_last_error()
    d.abi_version = _abi.ABI_VERSION
    d.m = 0
    assert mcb_lib.mcb200_layer_create(C.byref(d), C.byref(h)) == _abi.ERR_INVALID
    d.m = 10
    d.device = 10_000
    assert mcb_lib.mcb200_layer_create(C.byref(d), C.byref(h)) in (_abi.ERR_INVALID, _abi.ERR_CUDA)
    assert mcb_lib.mcb200_layer_simulate(None, -1, None) == _abi.ERR_INVALID
    assert mcb_lib.mcb200_layer_push(None, None, 3) == _abi.ERR_INVALID


def test_no_cpu_fallback(mcb_lib):
    """On a box without a CUDA device the product refuses to run instead of degrading."""
    from mc_mpi_b200 import _abi
    if mcb_lib.mcb200_device_count() > 0:
        pytest.skip("a GPU is present; the no-device failure path cannot be exercised")
    from mc_mpi_b200.layer import Layer
    with pytest.raises(_abi.McbError) as ei:
        Layer(0.0, 1.0, 0, 100, 0.0)
    assert ei.value.code in (_abi.ERR_CUDA, _abi.ERR_INVALID)
    x = np.ones(4, dtype=np.float32)
    y = np.empty_like(x)
    assert mcb_lib.mcb200_test_logf(0, x.ctypes.data, y.ctypes.data, 4) != _abi.OK


def test_product_never_imports_the_oracle():
    """oracle/ is checker-only: nothing under mc_mpi_b200/ or include/ may reference it."""
    bad = []
    for base in ("mc_mpi_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|pyoracle|mc_oracle|libmcref", txt):
                        # comments that merely NAME the checker are fine in docs, not in code paths
                        for line in txt.splitlines():
                            if re.search(r"(from|import)\s+oracle|pyoracle|dlopen.*oracle|libmcref", line):
                                bad.append((f, line.strip()))
    assert not bad, bad
