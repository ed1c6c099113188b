"""The C-ABI library loads and exports every symbol include/wavernn_b200.h declares.
No compute calls (no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

import helpers
from wavernn_b200 import cabi

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not cabi.is_built():
        import __graft_entry__ as g
        g.build()
    return cabi.load()


def test_header_symbols_are_exported(lib):
    header = (ROOT / "include" / "wavernn_b200.h").read_text()
    declared = set(re.findall(r"\b(wrnn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.wrnn_abi_version() == cabi.ABI_VERSION == 5


def test_struct_layouts_match_header():
    assert ctypes.sizeof(cabi.WrnnCfg) == 8 * 4
    assert ctypes.sizeof(cabi.WrnnWeights) == 16 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(cabi.WrnnJob) == 8 * 4 + 4 * 4 + 2 * 8 + 2 * 8 + 3 * 8 + 2 * 8 + 3 * 8 + 2 * 4 + 8      # + uniforms_ready (ABI v5)
    assert cabi.WrnnJob.seg_first.offset == 40 and cabi.WrnnJob.uniforms.offset == 48 and cabi.WrnnJob.out.offset == 80


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_create_without_device_reports_no_fallback(lib):
    m = helpers.make_model(0, "MOL")
    with pytest.raises(cabi.EngineError) as e:
        cabi.Engine(m.hot_state())
    assert e.value.code == cabi.WRNN_E_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_bad_arguments_are_rejected(lib):
    h = ctypes.c_void_p()
    assert lib.wrnn_create(ctypes.byref(h), None, None, 0) == cabi.WRNN_E_INVALID
    cfg = cabi.WrnnCfg(256, 512, 80, 32, 30, 0, 0, 0)
    w = cabi.WrnnWeights()
    assert lib.wrnn_create(ctypes.byref(h), ctypes.byref(cfg), ctypes.byref(w), 0) == cabi.WRNN_E_INVALID
    assert b"unsupported dims" in lib.wrnn_last_error()
    assert lib.wrnn_generate(None, None, None) == cabi.WRNN_E_INVALID
    assert lib.wrnn_check(None) == cabi.WRNN_E_INVALID
