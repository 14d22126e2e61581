"""CPU tests: the oracle (oracle/gunrock_oracle.c) against the golden vectors minted from the
unmodified reference (tests/golden/make_golden.py) and, where oracle/_ref exists, against the
compiled reference itself."""
import hashlib
import os
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT, bits_to_f32

INT_MAX = 2**31 - 1


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def write_mtx(path, header, n, entries):
    with open(path, "w") as f:
        f.write(header + "\n% comment line\n")
        f.write(f"{n} {n} {len(entries)}\n")
        for e in entries:
            f.write(" ".join(str(x) for x in e) + "\n")


def test_chesapeake_loader_and_csr(golden, tmp_path):
    g = golden["chesapeake"]
    # re-create the .mtx from the golden COO (the dataset itself does not travel to the GPU box)
    I, J = np.array(g["coo_I"]), np.array(g["coo_J"])
    # the symmetric loader mirrors every off-diagonal entry in place: originals are the even slots
    entries = [(int(I[k]) + 1, int(J[k]) + 1) for k in range(0, len(I), 2)]
    p = tmp_path / "ches.mtx"
    write_mtx(p, "%%MatrixMarket matrix coordinate pattern symmetric", g["n_rows"], entries)
    m = oracle.load_mtx(str(p))
    assert m["nnz"] == g["nnz"] == 340 and m["n_rows"] == 39
    assert (m["directed"], m["weighted"], m["symmetric"]) == (g["directed"], g["weighted"], g["symmetric"])
    assert np.array_equal(m["I"], I) and np.array_equal(m["J"], J)
    ro, ci, v = oracle.csr_from_coo(m["n_rows"], m["I"], m["J"], m["V"])
    assert ro.tolist() == g["row_offsets"] and ci.tolist() == g["column_indices"]
    assert np.array_equal(v, bits_to_f32(g["values_bits"]))
    # SURVEY.md 8c golden head
    assert ro[:8].tolist() == [0, 11, 22, 29, 33, 37, 41, 51]
    assert ci[:10].tolist() == [6, 7, 10, 11, 12, 21, 22, 33, 34, 36]


@pytest.mark.parametrize("name", ["chesapeake", "pytest_dag", "sample4"])
def test_bfs_sssp_small_golden(golden, name):
    g = golden[name]
    ro = np.array(g["row_offsets"], np.int32)
    ci = np.array(g["column_indices"], np.int32)
    v = bits_to_f32(g.get("values_bits", g.get("V_bits")))
    if name == "pytest_dag":
        ro, ci, v = oracle.csr_from_coo(5, np.array(g["I"]), np.array(g["J"]), v)
        assert ro.tolist() == g["row_offsets"] and ci.tolist() == g["column_indices"]
    for s, exp in g["bfs"].items():
        assert oracle.bfs(ro, ci, int(s)).tolist() == exp
    for s, exp in g["sssp"].items():
        got = oracle.sssp(ro, ci, v, int(s))
        assert np.array_equal(got.view(np.uint32), np.array(exp, np.uint32))


def test_known_answers_from_survey(golden):
    assert golden["chesapeake"]["bfs"]["0"] == [0, 2, 2, 2, 2, 2, 1, 1, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2,
                                                 2, 2, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 2, 1, 2, 1]
    assert golden["pytest_dag"]["bfs"]["0"] == [0, 1, 1, 2, 3]
    assert bits_to_f32(golden["pytest_dag"]["sssp"]["0"]).tolist() == [0.0, 1.0, 2.0, 2.5, 5.0]


@pytest.mark.parametrize("name", ["rmat10", "rmat12w"])
def test_rmat_golden(golden, name):
    g = golden[name]
    ro, ci = oracle.rmat_csr(g["scale"], g["edge_factor"], g["seed"])
    assert len(ci) == g["nnz"] and sha(ro, ci) == g["csr_sha256"]
    w = oracle.edge_weights(g["weight_seed"], ro, ci, g["non_integer"])
    assert sha(w) == g["weights_sha256"]
    for s in g["sources"]:
        assert oracle.bfs(ro, ci, s).tolist() == g["bfs"][str(s)]
        assert np.array_equal(oracle.sssp(ro, ci, w, s).view(np.uint32),
                              np.array(g["sssp"][str(s)], np.uint32))


def test_rmat_structure():
    ro, ci = oracle.rmat_csr(12, 16, 99)
    V = 1 << 12
    assert ro[0] == 0 and ro[-1] == len(ci) and np.all(np.diff(ro) >= 0)
    src = np.repeat(np.arange(V, dtype=np.int64), np.diff(ro))
    assert not np.any(src == ci)                           # no self loops
    keys = src * V + ci
    assert np.all(np.diff(keys) > 0)                       # sorted, deduplicated
    assert np.array_equal(np.sort(ci.astype(np.int64) * V + src), keys)  # symmetric
    w = oracle.edge_weights(5, ro, ci, False)
    assert w.min() >= 1 and w.max() <= 63 and np.all(w == np.round(w))
    # symmetric weights: w(u,v) == w(v,u)
    order = np.argsort(ci.astype(np.int64) * V + src, kind="stable")
    assert np.array_equal(w[order], w)


@pytest.mark.parametrize("scale,ef,seed,mirror,fold", [(12, 16, 5, True, 0), (14, 8, 7, True, 0),
                                                        (13, 17, 9, False, 5000), (10, 4, 3, True, 0), (4, 2, 1, True, 0)])
def test_parallel_host_build_equals_the_serial_one(scale, ef, seed, mirror, fold):
    """The OpenMP builder the bench's reference arm uses at RMAT-24/26 (orc_rmat_csr_parallel) gives the same
    CSR, bit for bit, as rmat_edges (+ fold) + build_csr_from_pairs -- which the device generator is pinned to."""
    n = ef * (1 << scale)
    s, d = oracle.rmat_edges(scale, n, seed)
    V = fold or (1 << scale)
    if fold:
        s, d = (s % fold).astype(np.int32), (d % fold).astype(np.int32)
    ro, ci = oracle.build_csr_from_pairs(V, s, d, mirror)
    ro2, ci2 = oracle.rmat_csr_parallel(scale, n, seed, mirror, fold)
    assert np.array_equal(ro, ro2) and np.array_equal(ci, ci2)
    for non_integer in (False, True):
        assert np.array_equal(oracle.edge_weights(3, ro, ci, non_integer),
                              oracle.edge_weights_parallel(3, ro, ci, non_integer))


def test_empty_and_degenerate():
    # single vertex, no edges
    ro = np.zeros(2, np.int32)
    ci = np.zeros(0, np.int32)
    assert oracle.bfs(ro, ci, 0).tolist() == [0]
    assert oracle.sssp(ro, ci, np.zeros(0, np.float32), 0).tolist() == [0.0]
    # unreachable vertices keep INT_MAX / FLT_MAX; self loops and duplicate edges are harmless
    I = np.array([0, 0, 0, 1, 1], np.int32)
    J = np.array([1, 1, 0, 1, 0], np.int32)
    V = np.array([2, 1, 7, 1, 4], np.float32)
    ro, ci, v = oracle.csr_from_coo(4, I, J, V)
    assert ro.tolist() == [0, 3, 5, 5, 5] and ci.tolist() == [1, 1, 0, 1, 0]  # stable, dups kept
    assert oracle.bfs(ro, ci, 0).tolist() == [0, 1, INT_MAX, INT_MAX]
    d = oracle.sssp(ro, ci, v, 0)
    assert d[:2].tolist() == [0.0, 1.0] and np.all(d[2:] == np.finfo(np.float32).max)


def test_mtx_loader_variants(tmp_path):
    p = tmp_path / "g.mtx"
    write_mtx(p, "%%MatrixMarket matrix coordinate real general", 3, [(1, 2, 0.5), (3, 1, 2.25), (2, 2, 7)])
    m = oracle.load_mtx(str(p))
    assert m["directed"] and m["weighted"] and not m["symmetric"]
    assert m["I"].tolist() == [0, 2, 1] and m["J"].tolist() == [1, 0, 1] and m["V"].tolist() == [0.5, 2.25, 7.0]
    write_mtx(p, "%%MatrixMarket matrix coordinate integer symmetric", 3, [(2, 1, 4), (3, 3, 9)])
    m = oracle.load_mtx(str(p))
    assert m["I"].tolist() == [1, 0, 2] and m["J"].tolist() == [0, 1, 2] and m["V"].tolist() == [4, 4, 9]
    (tmp_path / "bad.mtx").write_text("not a banner\n")
    with pytest.raises(RuntimeError):
        oracle.load_mtx(str(tmp_path / "bad.mtx"))


def test_transpose():
    ro, ci = oracle.rmat_csr(8, 4, 3, mirror=False)
    w = oracle.edge_weights(1, ro, ci, True)
    V = 1 << 8
    t_ro, t_ci, t_w = oracle.csr_transpose(V, V, ro, ci, w)
    src = np.repeat(np.arange(V), np.diff(ro))
    order = np.lexsort((src, ci))
    assert np.array_equal(t_ci, src[order]) and np.array_equal(t_w, w[order])
    assert np.array_equal(np.diff(t_ro), np.bincount(ci, minlength=V))


def test_pagerank_against_float64_power_iteration():
    """orc_pr is UNPINNED against the reference (it has no CPU PageRank); this checks
    it against an independent float64 power iteration of the same recurrence."""
    ro, ci = oracle.rmat_csr(10, 8, 77, mirror=False)
    V = 1 << 10
    p, iters = oracle.pr(ro, ci, None, 0.85, 1e-6)
    assert 2 <= iters < 200
    deg = np.diff(ro).astype(np.float64)
    src = np.repeat(np.arange(V), np.diff(ro))
    x = np.full(V, 1.0 / V)
    for _ in range(iters):
        iw = np.where(deg > 0, 0.85 / np.maximum(deg, 1), 0.0)
        dsum = 0.85 * x[deg == 0].sum()
        nxt = np.full(V, (1 - 0.85 + dsum) / V)
        np.add.at(nxt, ci, x[src] * iw[src])
        x = nxt
    assert abs(p.sum() - 1.0) < 1e-4
    assert np.allclose(p, x, rtol=2e-5, atol=1e-9)
    # weighted variant: ranks still sum to one and respond to the weights
    w = oracle.edge_weights(3, ro, ci, True)
    pw, _ = oracle.pr(ro, ci, w, 0.85, 1e-6)
    assert abs(pw.sum() - 1.0) < 1e-4 and not np.allclose(pw, p)
    # max_iter cap
    p1, it1 = oracle.pr(ro, ci, None, 0.85, 1e-6, max_iter=1)
    assert it1 == 1


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_restatement_matches_compiled_reference():
    rng = np.random.default_rng(5)
    for scale, ef, seed in ((9, 8, 1), (11, 16, 2), (13, 4, 3)):
        ro, ci = oracle.rmat_csr(scale, ef, seed)
        w = oracle.edge_weights(seed, ro, ci, True)
        g = oracle.RefGraph(ro, ci, w)
        for s in rng.integers(0, 1 << scale, 4):
            assert np.array_equal(oracle.bfs(ro, ci, int(s)), g.bfs(int(s))[0])
            assert np.array_equal(oracle.sssp(ro, ci, w, int(s)).view(np.uint32),
                                  g.sssp(int(s))[0].view(np.uint32))
    # from_coo parity on a random COO with duplicates and self loops
    n, nnz = 200, 3000
    I = rng.integers(0, n, nnz).astype(np.int32)
    J = rng.integers(0, n, nnz).astype(np.int32)
    V = rng.random(nnz).astype(np.float32)
    a = oracle.csr_from_coo(n, I, J, V)
    b = oracle.ref_csr_from_coo(n, n, I, J, V)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# ---- the reference's second vendored dataset: bips98_606.mtx (directed, real, negative values) ----------
@pytest.fixture(scope="module")
def bips98():
    return np.load(os.path.join(ROOT, "tests", "golden", "bips98_606.npz"))


def test_bips98_restatement_matches_the_reference_outputs(bips98):
    """tests/golden/bips98_606.npz was minted by the compiled reference (loader, from_coo, bfs_cpu, sssp_cpu
    on |value|); the C restatement must reproduce every stored depth / distance bit for bit."""
    ro, ci = bips98["row_offsets"], bips98["column_indices"]
    w = np.abs(bips98["values_bits"].view(np.float32))
    assert bips98["props"].tolist() == [1, 1, 0] and len(ro) == 7136 and len(ci) == 34738
    for s in bips98["sources"].tolist():
        assert np.array_equal(oracle.bfs(ro, ci, s), bips98[f"bfs_{s}"])
        assert np.array_equal(oracle.sssp(ro, ci, w, s).view(np.uint32), bips98[f"sssp_abs_bits_{s}"])
    d = bips98["bfs_0"]
    assert d[0] == 0 and (d == 2**31 - 1).any() and (d < 2**31 - 1).sum() > 1000    # directed: not all reachable


@pytest.mark.skipif(not os.path.exists("/root/reference/datasets/bips98_606/bips98_606.mtx"),
                    reason="the reference tree is only present in the authoring container")
def test_bips98_loader_restatements_on_the_real_file(bips98):
    """The oracle's loader / from_coo and the pygunrock-surface loader read the real file (header comments,
    `.987`-style reals, negative values, explicit diagonal) into exactly the reference's CSR."""
    path = "/root/reference/datasets/bips98_606/bips98_606.mtx"
    m = oracle.load_mtx(path)
    ro, ci, v = oracle.csr_from_coo(m["n_rows"], m["I"], m["J"], m["V"])
    assert np.array_equal(ro, bips98["row_offsets"]) and np.array_equal(ci, bips98["column_indices"])
    assert np.array_equal(v.view(np.uint32), bips98["values_bits"])
    assert [int(m["directed"]), int(m["weighted"]), int(m["symmetric"])] == bips98["props"].tolist()
    sys.path.insert(0, os.path.join(ROOT, "python"))
    import gunrock
    props, coo = gunrock.matrix_market_t().load(path)
    csr = gunrock.csr_t()
    csr.from_coo(coo)
    assert np.array_equal(csr.row_offsets, ro) and np.array_equal(csr.column_indices, ci)
    assert np.array_equal(csr.nonzero_values.view(np.uint32), v.view(np.uint32))
    assert (props.directed, props.weighted, props.symmetric) == (True, True, False)
