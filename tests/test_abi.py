"""The C-ABI library loads and exports every symbol include/micloc_b200.h declares.
CPU only: no compute call is made."""
import ctypes
import os
import re

import pytest

import helpers as H
from haghighatshoarmuir2024_b200 import _native as N

HEADER = os.path.join(H.ROOT, "include", "micloc_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(micloc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("micloc_snn_create", "micloc_snn_run", "micloc_snn_run_taps", "micloc_snn_run_host",
                 "micloc_rzcc_encode_f64", "micloc_hilbert_beamform", "micloc_xylo_create", "micloc_xylo_run", "micloc_xylo_process",
                 "micloc_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(N.LIB_PATH):
        N.build()
    lib = ctypes.CDLL(N.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    assert sorted(N.SYMBOLS) == declared_functions()


def test_version_and_error_string_without_gpu():
    lib = N.lib()
    assert lib.micloc_version() == 100
    assert isinstance(lib.micloc_last_error(), bytes)
    assert lib.micloc_launch_count() >= 0


def test_bad_config_is_rejected_before_touching_the_gpu():
    lib = N.lib()
    cfg = N.SnnConfig()          # num_mic == 0
    h = ctypes.c_void_p()
    rc = lib.micloc_snn_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == N.ERR_CONFIG and b"num_mic" in lib.micloc_last_error()
    with pytest.raises(ValueError):
        N.check(rc)


def test_bad_xylo_config_is_rejected_before_touching_the_gpu():
    lib = N.lib()
    cfg = N.XyloConfig()         # num_mic == 0
    h = ctypes.c_void_p()
    rc = lib.micloc_xylo_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc == N.ERR_CONFIG and b"num_mic" in lib.micloc_last_error()
