"""CPU tests of the C-ABI library: it loads, exports every symbol include/ysm.h declares, its
host-side pieces agree with the oracle, and it fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle
from yag_slam_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ysm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ysm_[a-z_0-9]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _capi.lib()
    names = _declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert sorted(_capi.EXPORTS) == names


def _header_struct_fields(name):
    """Field names of `typedef struct name { ... } name;` in include/ysm.h, in declaration order."""
    src = open(os.path.join(ROOT, "include", "ysm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
        for n in names.split(","):
            out.append(re.sub(r"[\s\*]|\[.*?\]", "", n))
    return out


def test_struct_layouts_match_header():
    assert C.sizeof(_capi.YsmParams) == 11 * 8 + 4 + 4 + 8 + 4 + 4
    assert C.sizeof(_capi.YsmDims) == 10 * 4 + 8
    assert C.sizeof(_capi.YsmBatch) == 4 + 4 + 8 + 7 * 8 + 4 * 4 + 8 + 8
    assert _capi.RESULT_DTYPE.itemsize == 128
    for cls, name in ((_capi.YsmParams, "ysm_params"), (_capi.YsmBatch, "ysm_batch"), (_capi.YsmDims, "ysm_dims"),
                      (_capi.YsmOccScans, "ysm_occ_scans"), (_capi.YsmOccInfo, "ysm_occ_info"),
                      (_capi.YsmChainQuery, "ysm_chain_query")):
        assert [f[0] for f in cls._fields_] == _header_struct_fields(name), name
    assert list(_capi.RESULT_DTYPE.names) == _header_struct_fields("ysm_result")


def test_integration_md_stub_structs_match_header():
    """The ctypes stub a maintainer would paste from INTEGRATION.md declares the same fields, in the same
    order, as include/ysm.h (a stale stub passes a short struct and the library reads past it)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for cls, name in (("_Params", "ysm_params"), ("_Batch", "ysm_batch"), ("_OccScans", "ysm_occ_scans"),
                      ("_OccInfo", "ysm_occ_info")):
        body = re.search(r"class %s\(C\.Structure\):.*?\n\n" % cls, doc, flags=re.S).group(0)
        assert re.findall(r'"([a-z_0-9]+)"', body) == _header_struct_fields(name), cls
    assert re.findall(r'\("([a-z_]+)", "<', re.search(r"_RESULT = np\.dtype\(\[.*?\]\)", doc, flags=re.S).group(0)) \
        == _header_struct_fields("ysm_result")


def test_point_readings_bit_exact_vs_oracle():
    rng = np.random.default_rng(5)
    for _ in range(20):
        n = int(rng.integers(1, 800))
        r = rng.uniform(0.0, 25.0, n)
        args = (rng.uniform(-3.2, 0), rng.uniform(0.001, 0.02), 0.05, 20.0,
                rng.uniform(-30, 30), rng.uniform(-30, 30), rng.uniform(-3.1, 3.1))
        a = _capi.point_readings(r, *args)
        b = oracle.point_readings(r, *args)
        assert a.shape == b.shape and (a.view(np.uint64) == b.view(np.uint64)).all()
    assert _capi.point_readings(np.zeros(0), -1, 0.01, 0.05, 20, 0, 0, 0).shape == (0, 2)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from yag_slam_b200.matcher import ScanMatcherB200
    with pytest.raises(RuntimeError) as e:
        ScanMatcherB200()
    assert "CUDA" in str(e.value)
    from yag_slam_b200 import raytracing
    with pytest.raises(RuntimeError):
        raytracing.run_raytracing_sweep(np.full((10, 10), 255, np.uint8), [0.0], 5, 5)


def test_create_rejects_bad_parameters():
    p = _capi.YsmParams()
    h = C.c_void_p()
    assert _capi.lib().ysm_create(C.byref(p), 0, C.byref(h)) == _capi.YSM_EINVAL  # all zeros
    assert "invalid" in _capi.last_error(None)
    assert _capi.lib().ysm_create(None, 0, C.byref(h)) == _capi.YSM_EINVAL


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yag-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "karto_oracle" not in txt and "libkarto_oracle" not in txt, f
