"""CPU: the C-ABI shared library builds (nvcc cross-compiles without a GPU), loads, and exports every
symbol that include/dvf_b200.h declares; argument validation works without touching a device."""
import ctypes as C
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(REPO, "include", "dvf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvf_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from dvf_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from dvf_b200 import _lib
    names = header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dvf_b200.h but not exported by libdvf_b200.so"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding table and header disagree"


def test_version_and_strerror(lib):
    assert lib.dvf_version() == 2
    assert lib.dvf_strerror(0) == b"ok"
    assert b"workspace" in lib.dvf_strerror(-6)
    assert b"unknown" in lib.dvf_strerror(-99)


def test_argument_validation_needs_no_device(lib):
    from dvf_b200._lib import dvf_desc, dvf_level, dvf_loss_desc
    d = dvf_desc(0, 3, 8, 8, 0, 0, 0, 0)
    assert lib.dvf_inverse_warp_fwd(C.byref(d), None, None, None, None, None, None, None) == -1    # shape
    d = dvf_desc(1, 3, 8, 8, 0, 0, 7, 0)
    assert lib.dvf_inverse_warp_fwd(C.byref(d), None, None, None, None, None, None, None) == -2    # padding enum
    d = dvf_desc(1, 3, 8, 8, 1, 1, 0, 0)
    assert lib.dvf_inverse_warp_fwd(C.byref(d), None, None, None, None, None, None, None) == -5    # bf16/NHWC: not in this entry
    d = dvf_desc(1, 3, 8, 8, 0, 0, 0, 0)
    assert lib.dvf_inverse_warp_fwd(C.byref(d), None, None, None, None, None, None, None) == -4    # NULL
    assert lib.dvf_inverse_warp_bwd_workspace_bytes(C.byref(d)) > 0
    ld = dvf_loss_desc(2, 3, 5, 1, 0, 0, 0, 0)
    lv = (dvf_level * 1)()
    lv[0].H, lv[0].W = 8, 8
    assert lib.dvf_photo_loss_workspace_bytes(C.byref(ld), lv) == 0                                  # V > DVF_MAX_VIEWS
    ld = dvf_loss_desc(2, 3, 2, 1, 0, 0, 0, 0)
    n = lib.dvf_photo_loss_workspace_bytes(C.byref(ld), lv)
    assert n > 0
    assert lib.dvf_photo_loss_fused(C.byref(ld), lv, None, None, 0, None) == -4
    assert lib.dvf_pose_proj_fwd(None, None, None, 1, 1, 0, None, 0, None, None, None, None) == -4


def test_library_does_not_link_torch_or_the_oracle():
    from dvf_b200 import _lib
    import subprocess
    out = subprocess.run(["ldd", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "c10" not in out
