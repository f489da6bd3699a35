"""CPU checks on the drop-in boundary: libdashing_b200.so loads, exports every symbol include/dashing_b200.h
declares, and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dashing_b200.h")).read()
    return sorted(set(re.findall(r"DB200_API\s+[\w\s\*]+?\b(db200_\w+)\s*\(", hdr)))


def test_header_symbols_exported(capi):
    syms = declared_symbols()
    assert len(syms) >= 25
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/dashing_b200.h but not exported: {missing}"
    assert set(syms) == set(capi.EXPORTS), "capi.py binds a different symbol set than the header declares"
    assert capi.lib.db200_version() == 101


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under dashing_b200/ may import, link or call it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "dashing_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".c")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.replace("dm::parallel_fill oracle", ""), f"{f} mentions the oracle"


def test_no_cpu_fallback_without_device(capi):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is visible; this test documents the no-GPU behaviour")
    regs = np.zeros((4, 1024), dtype=np.uint8)
    for call in (lambda: capi.cardinalities(regs, 10),
                 lambda: capi.dist_symmetric(regs, 10),
                 lambda: capi.dist_rect(regs[:2], regs[2:], 10),
                 lambda: capi.sketch_genomes([b"ACGT" * 100], 31, 10),
                 lambda: capi.Sketcher(10, 31),
                 lambda: capi.DistPlan(0)):
        with pytest.raises(capi.Db200Error) as ei:
            call()
        assert ei.value.code == capi.ENODEV
        assert "no CPU fallback" in str(ei.value)


def test_argument_validation(capi):
    # argument checks come before the device check so hosts get a precise reason to keep the reference path
    with pytest.raises(capi.Db200Error) as ei:
        capi.sketch_genomes([b"ACGT" * 100], 33, 10)       # k > 32: reference refuses too (distmain.cpp:101-102)
    assert ei.value.code in (capi.EUNSUPPORTED, capi.ENODEV)
    with pytest.raises(capi.Db200Error) as ei:
        capi.Sketcher(10, 0)
    assert ei.value.code == capi.EUNSUPPORTED
