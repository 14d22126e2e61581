"""GPU parity tests: the CUDA path, called through the C ABI (gunrock_b200 -> libgunrock_b200.so),
against the oracle and the golden vectors minted from the reference.  Bit-exact for BFS depths and
SSSP fp32 distances, 1e-6 relative for PageRank (BASELINE.json north_star)."""
import numpy as np
import pytest

import oracle
from conftest import bits_to_f32

pytestmark = pytest.mark.gpu

INT_MAX = 2**31 - 1
FLT_MAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def gb(built):
    import gunrock_b200 as gb
    if gb.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    return gb


def all_option_sets(gb):
    LB, DIR = gb.load_balance_t, gb.advance_direction_t
    out = []
    for lb in (LB.thread_mapped, LB.block_mapped, LB.merge_path):
        for d in (DIR.forward, DIR.optimized, DIR.backward):
            out.append(gb.options_t(advance_load_balance=lb, advance_direction=d, hub_threshold=64))
    out.append(gb.options_t(hub_threshold=32, ctas_per_sm=1))
    out.append(gb.options_t(reference_functor=True))
    out.append(gb.options_t(advance_load_balance=LB.merge_path, reference_functor=True))
    return out


def golden_graph(golden, name):
    g = golden[name]
    if name.startswith("rmat"):
        ro, ci = oracle.rmat_csr(g["scale"], g["edge_factor"], g["seed"])
        w = oracle.edge_weights(g["weight_seed"], ro, ci, g["non_integer"])
        return ro, ci, w, True
    ro = np.array(g["row_offsets"], np.int32)
    ci = np.array(g["column_indices"], np.int32)
    v = bits_to_f32(g.get("values_bits", g.get("V_bits")))
    return ro, ci, v, name == "chesapeake"


@pytest.mark.parametrize("name", ["chesapeake", "pytest_dag", "sample4", "rmat10", "rmat12w"])
def test_bfs_golden(gb, golden, name):
    ro, ci, w, sym = golden_graph(golden, name)
    G = gb.graph_t.from_csr(ro, ci, w, symmetric=sym)
    for opt in all_option_sets(gb):
        for s, exp in golden[name]["bfs"].items():
            d = np.full(G.n_vertices, -7, np.int32)
            st = gb.bfs(G, int(s), d, options=opt)
            assert d.tolist() == exp, (name, s, opt)
            assert st.kernel_launches > 0 and st.iterations >= 1
    G.close()


@pytest.mark.parametrize("name", ["chesapeake", "pytest_dag", "sample4", "rmat10", "rmat12w"])
def test_sssp_golden_bit_exact(gb, golden, name):
    ro, ci, w, sym = golden_graph(golden, name)
    G = gb.graph_t.from_csr(ro, ci, w, symmetric=sym)
    LB = gb.load_balance_t
    for lb in (LB.thread_mapped, LB.block_mapped, LB.merge_path):
        for s, exp in golden[name]["sssp"].items():
            d = np.zeros(G.n_vertices, np.float32)
            gb.sssp(G, int(s), d, options=gb.options_t(advance_load_balance=lb, hub_threshold=64))
            assert np.array_equal(d.view(np.uint32), np.array(exp, np.uint32)), (name, s, lb)
    G.close()


@pytest.mark.parametrize("scale,ef,seed", [(14, 16, 1), (16, 16, 0x5EED22), (17, 8, 5)])
def test_bfs_sssp_vs_oracle_medium(gb, scale, ef, seed):
    ro, ci = oracle.rmat_csr(scale, ef, seed)
    w = oracle.edge_weights(seed, ro, ci, True)
    G = gb.graph_t.from_csr(ro, ci, w, symmetric=True)
    deg = np.diff(ro)
    rng = np.random.default_rng(seed)
    sources = [int(deg.argmax()), int(rng.choice(np.flatnonzero(deg > 0))), int(np.flatnonzero(deg == 0)[0])]
    LB, DIR = gb.load_balance_t, gb.advance_direction_t
    for s in sources:
        exp = oracle.bfs(ro, ci, s)
        for lb in (LB.block_mapped, LB.merge_path):
            for d_ in (DIR.forward, DIR.optimized):
                for hub in (128, 4096):
                    d = np.empty(G.n_vertices, np.int32)
                    st = gb.bfs(G, s, d, options=gb.options_t(advance_load_balance=lb, advance_direction=d_,
                                                              hub_threshold=hub))
                    assert np.array_equal(d, exp), (s, lb, d_, hub)
                    if d_ == DIR.forward:  # push touches every out-edge of every reached vertex once
                        assert st.edges_touched == int(deg[exp < INT_MAX].sum())
        exp_s = oracle.sssp(ro, ci, w, s)
        for lb in (LB.block_mapped, LB.merge_path):
            d = np.empty(G.n_vertices, np.float32)
            gb.sssp(G, s, d, options=gb.options_t(advance_load_balance=lb, hub_threshold=256))
            assert np.array_equal(d.view(np.uint32), exp_s.view(np.uint32)), (s, lb)
    G.close()


def test_directed_graph_pull_uses_transpose(gb):
    ro, ci = oracle.rmat_csr(13, 8, 21, mirror=False)
    G = gb.graph_t.from_csr(ro, ci, None, symmetric=False)
    s = int(np.diff(ro).argmax())
    exp = oracle.bfs(ro, ci, s)
    for d_ in (gb.advance_direction_t.optimized, gb.advance_direction_t.backward):
        d = np.empty(G.n_vertices, np.int32)
        st = gb.bfs(G, s, d, options=gb.options_t(advance_direction=d_))
        assert np.array_equal(d, exp)
        assert 1 in st.level_direction
    G.close()


def test_edge_cases(gb):
    # one vertex, no edges
    G = gb.graph_t.from_csr(np.zeros(2, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32))
    d = np.empty(1, np.int32)
    gb.bfs(G, 0, d)
    assert d.tolist() == [0]
    f = np.empty(1, np.float32)
    gb.sssp(G, 0, f)
    assert f.tolist() == [0.0]
    with pytest.raises(gb.GunrockB200Error):
        gb.bfs(G, 5, d)
    G.close()
    # duplicates, self loops, unreachable tail, zero-degree source
    I = np.array([0, 0, 0, 1, 1, 3], np.int32)
    J = np.array([1, 1, 0, 1, 0, 0], np.int32)
    V = np.array([2, 1, 7, 1, 4, 3], np.float32)
    ro, ci, v = oracle.csr_from_coo(5, I, J, V)
    G = gb.graph_t.from_csr(ro, ci, v)
    for s in range(5):
        d = np.empty(5, np.int32)
        gb.bfs(G, s, d)
        assert np.array_equal(d, oracle.bfs(ro, ci, s))
        f = np.empty(5, np.float32)
        gb.sssp(G, s, f)
        assert np.array_equal(f.view(np.uint32), oracle.sssp(ro, ci, v, s).view(np.uint32))
    G.close()
    # a star: one hub row far above the hub threshold, forces the TMA slab bin, ragged tail sizes
    for n in (33, 2049, 4097, 10001):
        ro = np.concatenate([[0], np.full(n, n - 1)]).astype(np.int32)
        ro[1:] = n - 1
        ci = np.arange(1, n, dtype=np.int32)
        G = gb.graph_t.from_csr(ro, ci, np.ones(n - 1, np.float32))
        d = np.empty(n, np.int32)
        gb.bfs(G, 0, d, options=gb.options_t(hub_threshold=32))
        assert d[0] == 0 and np.all(d[1:] == 1)
        G.close()


def test_many_deferred_rows_and_small_tickets(gb):
    """Every vertex has 40 out-edges and hub_threshold is 32, so whole frontiers (hundreds of
    thousands of rows) are deferred to the TMA slab kernel: exercises the 'more batches than CTAs'
    ownership path of advance_hub_kernel, and the CTA-scan kernel with tiny / full tickets."""
    rng = np.random.default_rng(2024)
    V, d = 600_000, 40
    ci = rng.integers(0, V, V * d).astype(np.int32)
    ro = (np.arange(V + 1, dtype=np.int64) * d).astype(np.int32)
    ci = np.sort(ci.reshape(V, d), axis=1).reshape(-1)          # sorted rows (duplicates allowed)
    w = (1.0 + rng.random(V * d)).astype(np.float32)
    G = gb.graph_t.from_csr(ro, ci, w, symmetric=False)
    exp = oracle.bfs(ro, ci, 7)
    exp_s = oracle.sssp(ro, ci, w, 7)
    for hub in (32, 4096):
        for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
            dd = np.empty(V, np.int32)
            st = gb.bfs(G, 7, dd, options=gb.options_t(advance_load_balance=lb, hub_threshold=hub))
            assert np.array_equal(dd, exp), (hub, lb)
            assert st.edges_touched == int((exp < INT_MAX).sum()) * d
        f = np.empty(V, np.float32)
        gb.sssp(G, 7, f, options=gb.options_t(hub_threshold=hub))
        assert np.array_equal(f.view(np.uint32), exp_s.view(np.uint32)), hub
    G.close()


def test_device_tensor_results_and_stream(gb):
    import torch
    ro, ci = oracle.rmat_csr(12, 16, 9)
    w = oracle.edge_weights(9, ro, ci, False)
    tro, tci, tw = (torch.from_numpy(x).cuda() for x in (ro, ci, w))
    G = gb.graph_t.view_csr(tro, tci, tw, symmetric=True)   # non-owning, graph_t semantics
    s = int(np.diff(ro).argmax())
    d = torch.empty(G.n_vertices, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        gb.bfs(G, s, d, options=gb.options_t(stream=stream.cuda_stream))
    stream.synchronize()
    assert np.array_equal(d.cpu().numpy(), oracle.bfs(ro, ci, s))
    f = torch.empty(G.n_vertices, dtype=torch.float32, device="cuda")
    gb.sssp(G, s, f)
    assert np.array_equal(f.cpu().numpy().view(np.uint32), oracle.sssp(ro, ci, w, s).view(np.uint32))
    G.close()


@pytest.mark.parametrize("tail", [1, 2, 3])
def test_views_over_unpadded_arrays_whose_last_row_is_a_hub(gb, tail):
    """A non-owning view (b2g_graph_view_csr) over arrays of EXACTLY E elements with E % 4 != 0 whose last row is
    a hub: the TMA slab copies of the hub bin and of the PageRank tiles move whole 16-byte groups and must stop at
    the end of the caller's arrays -- the trailing 1-3 elements arrive through plain loads.  The arrays sit at the
    very end of one exactly-sized allocation, sentinels behind them would show a read past the end as a wrong
    neighbour."""
    import torch
    n = 9000
    hub = n - 1
    deg_hub = 6002 + tail                  # directed star into a chain: hub -> 0..deg_hub-1, v -> v+1
    src = np.concatenate([np.arange(n - 2), np.full(deg_hub, hub)]).astype(np.int64)
    dst = np.concatenate([np.arange(1, n - 1), np.arange(deg_hub)]).astype(np.int64)
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    ro = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(src, minlength=n), out=ro[1:])
    ci = dst.astype(np.int32)
    E = len(ci)
    assert E % 4 == tail
    w = oracle.edge_weights(3, ro, ci, True)
    # one allocation: [ro | ci | w | poison]: ci and w end exactly where the next array / the poison begins
    pad_ro = (-(n + 1)) % 4
    blob = torch.full((n + 1 + pad_ro + E + (-E) % 4 + E + 64,), -1, dtype=torch.int32, device="cuda")
    o_ci = n + 1 + pad_ro
    o_w = o_ci + E + (-E) % 4
    blob[:n + 1] = torch.from_numpy(ro).cuda()
    blob[o_ci:o_ci + E] = torch.from_numpy(ci).cuda()
    blob[o_ci + E:o_w] = 0x7fffffff                      # what an over-read of the column indices would see
    tw = blob[o_w:o_w + E].view(torch.float32)
    tw.copy_(torch.from_numpy(w).cuda())
    blob[o_w + E:] = 0x7fc00000                          # NaN bits behind the weights
    torch.cuda.synchronize()      # the fills ran on torch's stream; the library launches on its own non-blocking one
    G = gb.graph_t.view_csr(blob[:n + 1], blob[o_ci:o_ci + E], tw, symmetric=False)
    for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
        for thr in (256, 4096):
            opt = gb.options_t(advance_load_balance=lb, hub_threshold=thr)
            d = np.empty(n, np.int32)
            gb.bfs(G, hub, d, options=opt)
            assert np.array_equal(d, oracle.bfs(ro, ci, hub)), (lb, thr)
            f = np.empty(n, np.float32)
            gb.sssp(G, hub, f, options=opt)
            assert np.array_equal(f.view(np.uint32), oracle.sssp(ro, ci, w, hub).view(np.uint32)), (lb, thr)
    G.build_transpose()
    p = np.empty(n, np.float32)
    st = gb.pr(G, p)
    pe, it = oracle.pr(ro, ci, w)
    assert st.iterations == it and np.allclose(p, pe, rtol=1e-6, atol=0)
    G.close()


@pytest.mark.parametrize("mirror,weights", [(True, None), (False, None), (False, True)])
def test_pagerank_vs_oracle(gb, mirror, weights):
    ro, ci = oracle.rmat_csr(13, 8, 31, mirror=mirror)
    w = oracle.edge_weights(4, ro, ci, True) if weights else None
    G = gb.graph_t.from_csr(ro, ci, w, symmetric=mirror)
    p = np.empty(G.n_vertices, np.float32)
    st = gb.pr(G, p, alpha=0.85, tol=1e-6)
    pe, iters = oracle.pr(ro, ci, w, 0.85, 1e-6)
    assert st.iterations == iters
    rel = np.abs(p - pe) / np.maximum(np.abs(pe), np.finfo(np.float32).tiny)
    assert rel.max() <= 1e-6, rel.max()          # tolerance stated by BASELINE.json north_star
    assert abs(float(p.sum()) - 1.0) < 1e-3
    p1 = np.empty(G.n_vertices, np.float32)
    assert gb.pr(G, p1, max_iter=3).iterations == 3
    pe3, _ = oracle.pr(ro, ci, w, 0.85, 1e-6, max_iter=3)
    assert np.abs(p1 - pe3).max() <= 1e-6 * np.abs(pe3).max()
    G.close()


def test_pagerank_tile_edge_cases(gb):
    """Rows that span many pull tiles (hubs), empty rows at tile borders, graphs with no edges."""
    rng = np.random.default_rng(11)
    cases = []
    # no edges at all: every rank is (1 - alpha + alpha)/V = 1/V after one iteration
    cases.append((np.zeros(6, np.int32), np.zeros(0, np.int32)))
    # in-star: vertex 0 has in-degree 9000 (spans 5 tiles of 2048), everyone else dangling or leaf
    n = 9001
    I = np.arange(1, n, dtype=np.int32)
    J = np.zeros(n - 1, np.int32)
    cases.append(oracle.csr_from_coo(n, I, J)[:2])
    # two hubs whose in-edge ranges straddle tile boundaries + a block of isolated vertices between
    n = 12000
    I = np.concatenate([rng.integers(0, n, 3000), rng.integers(0, n, 5000), rng.integers(0, n, 6000)]).astype(np.int32)
    J = np.concatenate([np.full(3000, 5), np.full(5000, 7000), rng.integers(7001, n, 6000)]).astype(np.int32)
    cases.append(oracle.csr_from_coo(n, I, J)[:2])
    for ro, ci in cases:
        G = gb.graph_t.from_csr(ro, ci, None, symmetric=False)
        p = np.empty(G.n_vertices, np.float32)
        st = gb.pr(G, p)
        pe, iters = oracle.pr(ro, ci, None)
        assert st.iterations == iters
        assert np.allclose(p, pe, rtol=1e-6, atol=0), np.abs(p - pe).max()
        G.close()


def test_ingest_rmat_coo_transpose(gb):
    # device RMAT generator == oracle generator (same counter-based integer arithmetic)
    for scale, ef, seed, mirror in ((10, 16, 0x5EED10, True), (13, 8, 77, True), (12, 8, 5, False)):
        V = 1 << scale
        for wmode in (0, 1, 2):
            G = gb.graph_t.rmat(scale, ef * V, seed, mirror=mirror, weights=wmode, weight_seed=seed + 1)
            ro, ci, w = G.download()
            ero, eci = oracle.rmat_csr(scale, ef, seed, mirror=mirror)
            assert np.array_equal(ro, ero) and np.array_equal(ci, eci)
            if wmode:
                assert np.array_equal(w, oracle.edge_weights(seed + 1, ero, eci, wmode == 2))
            v, dmax = G.max_degree_vertex()
            deg = np.diff(ero)
            assert v == int(deg.argmax()) and dmax == int(deg.max())
            G.close()
    # COO -> CSR on the device == csr_t::from_coo (stable, duplicates kept)
    rng = np.random.default_rng(3)
    n, nnz = 300, 5000
    I = rng.integers(0, n, nnz).astype(np.int32)
    J = rng.integers(0, n, nnz).astype(np.int32)
    Vv = rng.random(nnz).astype(np.float32)
    G = gb.graph_t.from_coo(n, n, I, J, Vv)
    got = G.download()
    exp = oracle.csr_from_coo(n, I, J, Vv)
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)
    G.close()
    G = gb.graph_t.from_coo(4, 4, np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert G.download()[0].tolist() == [0, 0, 0, 0, 0]
    G.close()


def test_operators(gb):
    import torch
    ro, ci = oracle.rmat_csr(12, 8, 13)
    V = 1 << 12
    G = gb.graph_t.from_csr(ro, ci, None, symmetric=True)
    dev = "cuda"
    rng = np.random.default_rng(1)
    # ---- filter: predicated / remove / compact are stable selects; bypass marks -1
    fr = rng.integers(-1, V, 10000).astype(np.int32)
    mask = (rng.random(V) < 0.5).astype(np.uint8)
    t_in = torch.from_numpy(fr).to(dev)
    t_cnt = torch.tensor([len(fr)], dtype=torch.int32, device=dev)
    t_mask = torch.from_numpy(mask).to(dev)
    keep = (fr >= 0) & (mask[np.maximum(fr, 0)] != 0)
    for alg in (gb.filter_algorithm_t.predicated, gb.filter_algorithm_t.remove, gb.filter_algorithm_t.compact):
        t_out = torch.full((len(fr),), -5, dtype=torch.int32, device=dev)
        t_oc = torch.zeros(1, dtype=torch.int32, device=dev)
        gb.filter(G, alg, t_in, t_cnt, t_out, t_oc, t_mask)
        n = int(t_oc.item())
        assert n == int(keep.sum()) and np.array_equal(t_out[:n].cpu().numpy(), fr[keep])
    t_out = torch.empty(len(fr), dtype=torch.int32, device=dev)
    t_oc = torch.zeros(1, dtype=torch.int32, device=dev)
    gb.filter(G, gb.filter_algorithm_t.bypass, t_in, t_cnt, t_out, t_oc, t_mask)
    assert int(t_oc.item()) == len(fr)
    assert np.array_equal(t_out.cpu().numpy(), np.where(keep, fr, -1))
    # empty frontier
    t_cnt0 = torch.zeros(1, dtype=torch.int32, device=dev)
    t_oc.fill_(99)
    gb.filter(G, gb.filter_algorithm_t.predicated, t_in, t_cnt0, t_out, t_oc, None)
    assert int(t_oc.item()) == 0
    # ---- uniquify: best effort = adjacent unique; exact = sorted unique set (sort + unique)
    dup = np.repeat(rng.integers(0, V, 3000), rng.integers(1, 4, 3000)).astype(np.int32)
    t_in = torch.from_numpy(dup).to(dev)
    t_cnt = torch.tensor([len(dup)], dtype=torch.int32, device=dev)
    t_out = torch.empty(len(dup), dtype=torch.int32, device=dev)
    gb.uniquify(G, t_in, t_cnt, t_out, t_oc, best_effort=True)
    exp = dup[np.concatenate([[True], dup[1:] != dup[:-1]])]
    assert int(t_oc.item()) == len(exp) and np.array_equal(t_out[:len(exp)].cpu().numpy(), exp)
    for trial in range(2):  # twice: the bitmap must be left clean
        gb.uniquify(G, t_in, t_cnt, t_out, t_oc, best_effort=False)
        exp = np.unique(dup)
        assert int(t_oc.item()) == len(exp) and np.array_equal(t_out[:len(exp)].cpu().numpy(), exp)
    # ---- advance with the BFS claim functor: one level from a 2-vertex frontier
    deg = np.diff(ro)
    f0 = np.argsort(-deg)[:2].astype(np.int32)
    for lb in (gb.load_balance_t.thread_mapped, gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
        vis = np.zeros((V + 31) // 32, np.uint32)
        for v in f0:
            vis[v >> 5] |= np.uint32(1 << (v & 31))
        t_vis = torch.from_numpy(vis.view(np.int32)).to(dev)
        t_lab = torch.full((V,), INT_MAX, dtype=torch.int32, device=dev)
        t_f = torch.from_numpy(f0).to(dev)
        t_fc = torch.tensor([2], dtype=torch.int32, device=dev)
        t_o = torch.empty(V, dtype=torch.int32, device=dev)
        t_oc2 = torch.zeros(1, dtype=torch.int32, device=dev)
        e = gb.advance_bfs(G, t_f, t_fc, t_o, t_oc2, t_vis, t_lab, 1,
                           gb.options_t(advance_load_balance=lb, hub_threshold=64))
        nbrs = np.unique(np.concatenate([ci[ro[v]:ro[v + 1]] for v in f0]))
        nbrs = nbrs[~np.isin(nbrs, f0)]
        n = int(t_oc2.item())
        assert e == int(deg[f0].sum())
        assert np.array_equal(np.sort(t_o[:n].cpu().numpy()), nbrs)       # each claimed exactly once
        lab = t_lab.cpu().numpy()
        assert np.all(lab[nbrs] == 1) and (lab == 1).sum() == len(nbrs)
    G.close()


def test_full_size_properties_rmat20(gb):
    """Size-independent properties at a scale the oracle would take too long for inside the
    GPU suite: BFS depth consistency over every edge, SSSP triangle inequality + tightness."""
    import torch
    scale = 20
    G = gb.graph_t.rmat(scale, 16 << scale, 0x5EED22, mirror=True, weights=2, weight_seed=3)
    ro, ci, w = G.download()
    s, _ = G.max_degree_vertex()
    dev = "cuda"
    src = torch.repeat_interleave(torch.arange(G.n_vertices, device=dev),
                                  torch.from_numpy(np.diff(ro)).to(dev))
    dst = torch.from_numpy(ci).to(dev).long()
    for opt in (gb.options_t(), gb.options_t(advance_load_balance=gb.load_balance_t.merge_path),
                gb.options_t(advance_direction=gb.advance_direction_t.optimized)):
        d = torch.empty(G.n_vertices, dtype=torch.int32, device=dev)
        gb.bfs(G, s, d, options=opt)
        dl = d.long()
        reached = dl < INT_MAX
        assert int(dl[s]) == 0 and int((dl == 0).sum()) == 1
        du, dv = dl[src], dl[dst]
        assert bool(((du < INT_MAX) == (dv < INT_MAX)).all())            # symmetric graph: same component
        ok = reached[src]
        assert bool(((du[ok] - dv[ok]).abs() <= 1).all())                 # no edge skips a level
        best = torch.full((G.n_vertices,), INT_MAX, dtype=torch.long, device=dev)
        best.scatter_reduce_(0, dst[ok], du[ok], reduce="amin")
        inner = reached.clone()
        inner[s] = False
        assert bool((best[inner] == dl[inner] - 1).all())                 # every depth has a parent
    f = torch.empty(G.n_vertices, dtype=torch.float32, device=dev)
    gb.sssp(G, s, f)
    wt = torch.from_numpy(w).to(dev)
    reached = f < FLT_MAX
    ok = reached[src]
    cand = f[src][ok] + wt[ok]                                            # fp32 add, as the kernel
    assert bool((f[dst][ok] <= cand).all())                               # no relaxable edge left
    best = torch.full((G.n_vertices,), float("inf"), dtype=torch.float32, device=dev)
    best.scatter_reduce_(0, dst[ok], cand, reduce="amin")
    inner = reached.clone()
    inner[s] = False
    assert bool((best[inner] == f[inner]).all())                          # every distance is attained
    assert float(f[s]) == 0.0
    G.close()
