"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/gunrock_b200.h declares, and refuses (loudly) to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gunrock_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2g_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    import gunrock_b200 as gb
    L = gb.lib()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gunrock_b200.h but not exported"
    assert sorted(gb.exported_symbols()) == declared
    assert L.b2g_version() >= 100


def test_options_default_mirror_reference_defaults(built):
    import gunrock_b200 as gb
    o = gb._Options()
    gb.lib().b2g_options_default(C.byref(o))
    # include/gunrock/algorithms/algorithms.hxx:27-72
    assert o.advance_load_balance == gb.load_balance_t.block_mapped
    assert o.filter_algorithm == gb.filter_algorithm_t.predicated
    assert (o.enable_filter, o.enable_uniquify, o.best_effort_uniquify) == (0, 0, 1)
    assert o.uniquify_percent == 100.0
    d = gb.options_t()._c()
    for f, _ in gb._Options._fields_:
        assert getattr(d, f) == getattr(o, f), f


def test_no_silent_cpu_fallback(built):
    import gunrock_b200 as gb
    if gb.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    ro = np.array([0, 1, 1], np.int32)
    ci = np.array([1], np.int32)
    with pytest.raises(gb.GunrockB200Error, match="no CUDA device"):
        gb.graph_t.from_csr(ro, ci)
    with pytest.raises(gb.GunrockB200Error):
        gb.graph_t.rmat(4, 16, 1)


def test_bad_arguments_are_rejected(built):
    import gunrock_b200 as gb
    L = gb.lib()
    h = C.c_void_p()
    assert L.b2g_graph_create_csr(-1, 0, None, None, None, 0, 0, C.byref(h)) == -1
    assert b"bad arguments" in L.b2g_last_error()
    assert L.b2g_bfs(None, 0, None, None, 0, None) == -1
    assert L.b2g_graph_destroy(None) == 0


def test_product_never_imports_oracle():
    """The product path must not route through the checker."""
    for base, _, files in os.walk(os.path.join(ROOT, "gunrock_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hxx")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libgunrock_oracle" not in text and "orc_" not in text, f
    for top in ("include", "python"):
        for base, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".pyc"):
                    continue
                text = open(os.path.join(base, f)).read()
                assert "libgunrock_oracle" not in text and "orc_" not in text, f
                assert "import oracle" not in text and "from oracle" not in text, f
