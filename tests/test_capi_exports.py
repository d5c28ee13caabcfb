"""The C-ABI library builds, loads, exports every symbol include/cmax_b200.h declares, and fails
LOUDLY (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cmax_slam_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cmax_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmaxb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    declared = _declared()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_capi.EXPORTS) == declared   # the binding list mirrors the header exactly


def test_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_event_record_layout():
    from cmax_slam_b200 import synth
    dt = synth.EVENT_DTYPE
    assert dt.itemsize == 16
    assert [dt.fields[k][1] for k in ("x", "y", "sec", "nsec", "polarity")] == [0, 2, 4, 8, 12]


def test_version_and_error_string():
    L = _capi.lib()
    assert L.cmaxb_version() == 200
    assert isinstance(L.cmaxb_last_error(), bytes)
    assert L.cmaxb_kernel_name(1) == b"fe_scatter"


def test_no_cpu_fallback():
    """Without a device every create() must fail with CMAXB_ERR_CUDA -- never compute on the host."""
    if _capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from cmax_slam_b200 import synth
    from cmax_slam_b200.backend import EventWarperCMax
    from cmax_slam_b200.frontend import AngVelEstimatorCMax
    lut = synth.bearing_lut(16, 12, (10, 10, 8, 6))
    with pytest.raises(_capi.CmaxbError) as e:
        AngVelEstimatorCMax(16, 12, (10, 10, 8, 6), lut)
    assert e.value.code == -2
    with pytest.raises(_capi.CmaxbError) as e:
        EventWarperCMax(16, 12, lut, 64, 32)
    assert e.value.code == -2


def test_argument_validation_without_device():
    L = _capi.lib()
    h = C.c_void_p()
    assert L.cmaxb_fe_create(None, C.byref(h)) == -1
    cfg = _capi.FeCfg(2, 2, 1, 1, 0, 0, None, 1.0, 100, 0, 0, 0, None, 1)
    assert L.cmaxb_fe_create(C.byref(cfg), C.byref(h)) == -1
    assert b"configuration" in L.cmaxb_last_error()
    bc = _capi.BeCfg(16, 12, None, 64, 32, 1.0, 100, 1, 3, 0, 0, 0, None)
    assert L.cmaxb_be_create(C.byref(bc), C.byref(h)) == -1
