"""The reference's Python module surface (python/src/gunrock/bindings.cu, python/tests/*.py) served by
gunrock_b200.pygunrock over the C ABI.  Host-side pieces are checked on the CPU
against the oracle's pinned loader / from_coo; the algorithm calls are GPU tests that read like
python/tests/test_algorithms.py."""
import os
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "python"))
import gunrock  # noqa: E402  (python/gunrock -> gunrock_b200.pygunrock)


MTX_GENERAL_REAL = """%%MatrixMarket matrix coordinate real general
% the 5-vertex graph of python/tests/conftest.py
5 5 5
1 2 1.0
1 3 2.0
2 4 1.5
3 4 1.0
4 5 2.5
"""

MTX_SYMMETRIC_PATTERN = """%%MatrixMarket matrix coordinate pattern symmetric
%comment
%another
6 6 7
2 1
3 1
3 3
4 2
5 4
6 6
6 1
"""

MTX_INTEGER = """%%MatrixMarket matrix coordinate integer general
3 3 4
1 1 7
3 2 2
1 3 9
1 3 4
"""


@pytest.mark.parametrize("text", [MTX_GENERAL_REAL, MTX_SYMMETRIC_PATTERN, MTX_INTEGER])
def test_matrix_market_loader_and_from_coo_match_the_pinned_oracle(tmp_path, text):
    p = tmp_path / "g.mtx"
    p.write_text(text)
    props, coo = gunrock.matrix_market_t().load(str(p))
    m = oracle.load_mtx(str(p))          # pinned against the reference's loader (tests/test_oracle.py)
    assert (coo.number_of_rows, coo.number_of_columns, coo.number_of_nonzeros) == (m["n_rows"], m["n_cols"], m["nnz"])
    assert np.array_equal(coo.row_indices, m["I"]) and np.array_equal(coo.column_indices, m["J"])
    assert np.array_equal(coo.nonzero_values.view(np.uint32), m["V"].view(np.uint32))
    assert (props.directed, props.weighted, props.symmetric) == (m["directed"], m["weighted"], m["symmetric"])
    csr = gunrock.csr_t()
    csr.from_coo(coo)
    ro, ci, v = oracle.csr_from_coo(m["n_rows"], m["I"], m["J"], m["V"])
    assert np.array_equal(csr.row_offsets, ro) and np.array_equal(csr.column_indices, ci)
    assert np.array_equal(csr.nonzero_values.view(np.uint32), v.view(np.uint32))
    assert csr.number_of_nonzeros == m["nnz"] and csr.number_of_rows == m["n_rows"]


def test_loader_on_the_golden_chesapeake_entries(golden, tmp_path):
    g = golden["chesapeake"]
    I, J = np.array(g["coo_I"]), np.array(g["coo_J"])
    with open(tmp_path / "ches.mtx", "w") as f:     # originals are the even slots of the mirrored COO
        f.write("%%MatrixMarket matrix coordinate pattern symmetric\n39 39 170\n")
        for k in range(0, len(I), 2):
            f.write(f"{I[k] + 1} {J[k] + 1}\n")
    props, coo = gunrock.matrix_market_t().load(str(tmp_path / "ches.mtx"))
    assert coo.number_of_nonzeros == 340 and props.symmetric and not props.directed and not props.weighted
    assert coo.row_indices.tolist() == g["coo_I"] and coo.column_indices.tolist() == g["coo_J"]
    csr = gunrock.csr_t()
    csr.from_coo(coo)
    assert csr.row_offsets.tolist() == g["row_offsets"] and csr.column_indices.tolist() == g["column_indices"]


def test_loader_rejects_what_the_reference_rejects(tmp_path):
    bad = tmp_path / "bad.mtx"
    bad.write_text("%%MatrixMarket matrix array real general\n2 2\n1.0\n2.0\n3.0\n4.0\n")
    with pytest.raises(gunrock_error()):
        gunrock.matrix_market_t().load(str(bad))
    bad.write_text("not a banner\n")
    with pytest.raises(gunrock_error()):
        gunrock.matrix_market_t().load(str(bad))


def gunrock_error():
    from gunrock_b200 import GunrockB200Error
    return GunrockB200Error


def test_formats_enums_and_binary_round_trip(tmp_path):
    # python/tests/test_formats.py
    for cls in (gunrock.csr_t, gunrock.coo_t, gunrock.csc_t):
        assert cls() is not None
        x = cls(10, 10, 20)
        assert (x.number_of_rows, x.number_of_columns, x.number_of_nonzeros) == (10, 10, 20)
    assert hasattr(gunrock.memory_space_t, "host") and hasattr(gunrock.memory_space_t, "device")
    for name in ("csr", "csc", "coo", "invalid"):
        assert hasattr(gunrock.view_t, name)
    props = gunrock.graph_properties_t()          # python/tests/test_graph.py
    props.directed, props.weighted = True, True
    assert props.directed and props.weighted and not props.symmetric
    o = gunrock.options_t()
    o.enable_uniquify = True
    assert o.best_effort_uniquify is True and o.uniquify_percent == 100.0
    # .csr binary layout of csr.hxx:142-228
    ro, ci = oracle.rmat_csr(8, 4, 3)
    w = oracle.edge_weights(9, ro, ci, True)
    a = gunrock.csr_t(len(ro) - 1, len(ro) - 1, len(ci))
    a.row_offsets, a.column_indices, a.nonzero_values = ro, ci, w
    a.write_binary(str(tmp_path / "g.csr"))
    raw = np.fromfile(tmp_path / "g.csr", np.int32)
    assert raw[:3].tolist() == [len(ro) - 1, len(ro) - 1, len(ci)] and raw[3:3 + len(ro)].tolist() == ro.tolist()
    b = gunrock.csr_t()
    b.read_binary(str(tmp_path / "g.csr"))
    assert np.array_equal(b.row_offsets, ro) and np.array_equal(b.column_indices, ci)
    assert np.array_equal(b.nonzero_values.view(np.uint32), w.view(np.uint32))


# ---- GPU: python/tests/test_algorithms.py against this module ----------------------------------------
@pytest.fixture
def context(built):
    import gunrock_b200
    if gunrock_b200.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device")
    return gunrock.multi_context_t(0)


@pytest.fixture
def small_graph(tmp_path):
    p = tmp_path / "test_graph.mtx"
    p.write_text(MTX_GENERAL_REAL)
    properties, coo = gunrock.matrix_market_t().load(str(p))
    csr = gunrock.csr_t()
    csr.from_coo(coo)
    return gunrock.build_graph(properties, csr), properties


@pytest.mark.gpu
def test_sssp_and_bfs_basic(context, small_graph):
    import torch
    G, properties = small_graph
    n = G.get_number_of_vertices()
    assert n == 5 and G.get_number_of_edges() == 5
    distances = torch.full((n,), float("inf"), dtype=torch.float32, device="cuda")
    predecessors = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    elapsed = gunrock.sssp(G, 0, distances, predecessors, context)
    context.synchronize()
    assert elapsed > 0
    assert distances.cpu().tolist() == [0.0, 1.0, 2.0, 2.5, 5.0]        # 0->1->3 (2.5) beats 0->2->3 (3.0)
    hops = torch.full((n,), torch.iinfo(torch.int32).max, dtype=torch.int32, device="cuda")
    elapsed = gunrock.bfs(G, 0, hops, predecessors, context)
    context.synchronize()
    assert elapsed > 0 and hops.cpu().tolist() == [0, 1, 1, 2, 3]
    # options, tensor reuse, several sources (test_sssp_with_options / test_tensor_reuse / test_multiple_sources)
    options = gunrock.options_t()
    options.enable_uniquify = True
    for source in range(3):
        distances.fill_(float("inf"))
        assert gunrock.sssp(G, source, distances, predecessors, context, options) > 0
        context.synchronize()
        assert distances[source].item() == 0
        reachable = torch.isfinite(distances)
        assert reachable.sum().item() >= 1 and distances[reachable].min().item() == 0
    with pytest.raises(gunrock_error()):
        gunrock.sssp(G, 0, torch.zeros(n), predecessors, context)       # host tensor: rejected, no CPU path


@pytest.mark.gpu
def test_parity_through_the_python_module(context, tmp_path):
    """Same bar as tests/test_gpu_parity.py, through the reference-facing Python calls."""
    import torch
    ro, ci = oracle.rmat_csr(12, 8, 41)
    w = oracle.edge_weights(7, ro, ci, True)
    csr = gunrock.csr_t(len(ro) - 1, len(ro) - 1, len(ci))
    csr.row_offsets, csr.column_indices, csr.nonzero_values = ro, ci, w
    props = gunrock.graph_properties_t()
    props.weighted, props.symmetric = True, True
    G = gunrock.build_graph(props, csr)
    n = G.get_number_of_vertices()
    src = int(np.diff(ro).argmax())
    pred = torch.empty(n, dtype=torch.int32, device="cuda")
    for lb in (gunrock.load_balance_t.block_mapped, gunrock.load_balance_t.merge_path):
        o = gunrock.options_t()
        o.advance_load_balance = lb
        d = torch.empty(n, dtype=torch.int32, device="cuda")
        gunrock.bfs(G, src, d, pred, context, o)
        assert np.array_equal(d.cpu().numpy(), oracle.bfs(ro, ci, src))
        f = torch.empty(n, dtype=torch.float32, device="cuda")
        gunrock.sssp(G, src, f, pred, context, o)
        assert np.array_equal(f.cpu().numpy().view(np.uint32), oracle.sssp(ro, ci, w, src).view(np.uint32))
    p = torch.empty(n, dtype=torch.float32, device="cuda")
    assert gunrock.pr_run(G, gunrock.pr_param_t(0.85, 1e-6), gunrock.pr_result_t(p), context) > 0
    exp, _ = oracle.pr(ro, ci, w, 0.85, 1e-6)
    assert np.all(np.abs(p.cpu().numpy() - exp) <= 1e-6 * np.abs(exp))
    G.close()
