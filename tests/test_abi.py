"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/mmo_b200.h
declares, and refuses to compute (loudly, no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mmo_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mmo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = mmo_b200.lib()
    names = _declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/mmo_b200.h but not exported: {missing}"


def test_no_torch_types_in_the_header():
    src = open(os.path.join(ROOT, "include", "mmo_b200.h")).read()
    assert "torch" not in src.lower() and "at::" not in src and "std::" not in src


def test_build_info_names_the_arch():
    assert b"sm_100a" in mmo_b200.lib().mmo_build_info()


def _has_gpu():
    n = C.c_int(0)
    rc = mmo_b200.lib().mmo_device_count(C.byref(n))
    return rc == 0 and n.value > 0


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_init_fails_loudly_without_a_gpu():
    L = mmo_b200.lib()
    rc = L.mmo_init(C.c_int(0))
    assert rc == -2
    msg = L.mmo_last_error().decode()
    assert "no CPU fallback" in msg
    with pytest.raises(mmo_b200.MmoError):
        mmo_b200.init(0)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_calls_refuse_to_run_uninitialised():
    L = mmo_b200.lib()
    h = C.c_void_p()
    z = np.zeros(1)
    zp = z.ctypes.data_as(C.POINTER(C.c_double))
    a = np.array([6], np.int32)
    rc = L.mmo_receptor_create(C.c_int32(1), zp, zp, zp, zp, a.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(h))
    assert rc == -3 and h.value is None
    assert "no CPU fallback" in L.mmo_last_error().decode()


def test_topk_merge_is_host_side_and_deterministic():
    L = mmo_b200.lib()
    k = 4
    scores = np.array([[1.0, 2.0, 5.0, 0.0], [1.0, 0.5, 9.0, 9.0]])
    frames = np.array([[10, 20, 50, 0], [7, 5, 90, 91]], np.int64)
    counts = np.array([3, 2], np.int32)
    os_, of_ = np.empty(k), np.empty(k, np.int64)
    n = C.c_int32()
    rc = L.mmo_topk_merge(C.c_int32(2), C.c_int32(k), scores.ctypes.data_as(C.POINTER(C.c_double)),
                          frames.ctypes.data_as(C.POINTER(C.c_int64)), counts.ctypes.data_as(C.POINTER(C.c_int32)),
                          os_.ctypes.data_as(C.POINTER(C.c_double)), of_.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n))
    assert rc == 0 and n.value == 4
    assert os_.tolist() == [0.5, 1.0, 1.0, 2.0]
    assert of_.tolist() == [5, 7, 10, 20]          # tie on 1.0 goes to the smaller frame


def test_ocaml_stubs_parse_against_the_header_and_match_the_externals():
    """No OCaml toolchain in this image: the stubs cannot be compiled for real.  What CAN be checked: gcc parses
    and type-checks mmo_b200/ocaml/gpu_stubs.c against include/mmo_b200.h (stand-in caml/*.h under tests/fake_caml), and
    every `external` of gpu.ml names a primitive the stub file defines (native and bytecode entry)."""
    import re
    import subprocess
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                        "-Werror=int-conversion", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "fake_caml"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "mmo_b200", "ocaml", "gpu_stubs.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    ml = open(os.path.join(ROOT, "mmo_b200", "ocaml", "gpu.ml")).read()
    stubs = open(os.path.join(ROOT, "mmo_b200", "ocaml", "gpu_stubs.c")).read()
    defined = set(re.findall(r"CAMLprim value (mmo_ml_\w+)\(", stubs))
    named = set(re.findall(r'"(mmo_ml_\w+)"', ml))
    assert named and named <= defined, sorted(named - defined)
    assert "mmo_ml_mc_run" in named and "mmo_ml_scan" in named          # the simulate_lig external was missing
    # every library function a stub calls is declared in the header
    hdr = open(os.path.join(ROOT, "include", "mmo_b200.h")).read()
    for fn in set(re.findall(r"\b(mmo_[a-z0-9_]+)\(", stubs)) - defined:
        if fn.startswith("mmo_ml_"):
            continue
        assert re.search(r"\b" + fn + r"\(", hdr), fn


def test_integration_md_shows_the_current_ocaml_binding():
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_integration.py"), "--check"])
    assert r.returncode == 0, "INTEGRATION.md is stale: run python tools/gen_integration.py"
