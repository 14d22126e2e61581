"""The A/B switches the library keeps (all read from the environment, none needed in normal use):

  B2G_ADVANCE_VARIANT=0|1|4   which merge_path kernel serves the fused BFS / SSSP functors (default: the functor's
                              own choice -- warp-private spans, advance.cuh op_merge_path_kernel)
  B2G_BFS_PULL_LEGACY=1       the first-generation pull kernels of the direction-optimised BFS (one sweep / list
                              kernel per level) instead of the first-in-neighbour pair K1 + K2 (bfs.cuh)

Every choice must give the same bits.  (Round 1's other experiments -- on-chip copies of the visited map, 4096-edge
tiles, span prefetch, near/far SSSP, list-first pull -- were measured at the start of round 2, lost, and are gone:
profiles/r2_bench_matrix.md.)
"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = [pytest.mark.gpu]


# ---- the merge_path kernels (B2G_ADVANCE_VARIANT: 0 CTA tiles, 1 / 4 warp-private spans; read on every call) ---
@pytest.fixture
def variant_env():
    old = os.environ.get("B2G_ADVANCE_VARIANT")
    yield lambda v: os.environ.__setitem__("B2G_ADVANCE_VARIANT", str(v))
    if old is None:
        os.environ.pop("B2G_ADVANCE_VARIANT", None)
    else:
        os.environ["B2G_ADVANCE_VARIANT"] = old


@pytest.fixture(scope="module")
def level_graphs():
    """Host CSRs shared by every kernel choice (up to 2^23 vertices)."""
    import oracle
    return [(scale, *oracle.rmat_csr(scale, ef, seed))
            for scale, ef, seed in ((5, 4, 1), (12, 8, 13), (16, 16, 5), (21, 2, 99), (23, 1, 7))]


@pytest.mark.parametrize("variant", [0, 1, 4])
def test_merge_path_variants_one_level_exact(built, variant_env, level_graphs, variant):
    """One advance level with the BFS claim functor through b2g_advance_bfs, merge_path, for frontiers that
    stress the span staging: two hubs, every vertex (degree-0 rows included), duplicates, a single short
    row, high vertex ids.  Each unvisited neighbour is claimed exactly once."""
    import numpy as np
    import torch
    import oracle
    import gunrock_b200 as gb
    variant_env(variant)
    INT_MAX = 2**31 - 1
    rng = np.random.default_rng(variant)
    for scale, ro, ci in level_graphs:
        V = 1 << scale
        G = gb.graph_t.from_csr(ro, ci, None, symmetric=True)
        deg = np.diff(ro)
        frontiers = [np.argsort(-deg)[:2], np.arange(V), np.flatnonzero(deg == 0)[:100],
                     np.flatnonzero(deg == 1)[:1], np.repeat(np.argsort(-deg)[:50], 3),
                     rng.integers(V // 2, V, 5000)]
        for f0 in frontiers:
            f0 = f0.astype(np.int32)
            if len(f0) == 0:
                continue
            vis = np.zeros((V + 31) // 32 + 4, np.uint32)
            pre = rng.integers(0, V, V // 7)                             # some vertices already visited
            np.bitwise_or.at(vis, pre >> 5, (np.uint32(1) << (pre & 31).astype(np.uint32)))
            t_vis = torch.from_numpy(vis.view(np.int32).copy()).cuda()
            t_lab = torch.full((V,), INT_MAX, dtype=torch.int32, device="cuda")
            t_f = torch.from_numpy(f0).cuda()
            t_fc = torch.tensor([len(f0)], dtype=torch.int32, device="cuda")
            cap = int(min(deg[f0].astype(np.int64).sum(), V)) + 64
            t_o = torch.empty(cap, dtype=torch.int32, device="cuda")
            t_oc = torch.zeros(1, dtype=torch.int32, device="cuda")
            e = gb.advance_bfs(G, t_f, t_fc, t_o, t_oc, t_vis, t_lab, 7,
                               gb.options_t(advance_load_balance=gb.load_balance_t.merge_path))
            assert e == int(deg[f0].astype(np.int64).sum())
            in_f = np.zeros(V, bool)
            in_f[f0] = True
            nb = np.unique(ci[np.repeat(in_f, deg)]).astype(np.int64)   # neighbours of the frontier's rows
            was = (vis[nb >> 5] >> (nb & 31).astype(np.uint32)) & 1
            exp = nb[was == 0]
            n = int(t_oc.item())
            got = np.sort(t_o[:n].cpu().numpy())
            assert np.array_equal(got, exp), (scale, len(f0), n, len(exp))
            lab = t_lab.cpu().numpy()
            assert (lab == 7).sum() == len(exp) and np.all(lab[exp] == 7)
            vis2 = t_vis.cpu().numpy().view(np.uint32)
            np.bitwise_or.at(vis, exp >> 5, (np.uint32(1) << (exp & 31).astype(np.uint32)))
            assert np.array_equal(vis2, vis)
        G.close()


@pytest.mark.parametrize("variant", [0, 1, 4])
def test_merge_path_variants_full_runs_bit_exact(built, variant_env, variant):
    """Whole BFS / SSSP runs at a size where the enactors really pick merge_path (frontier out-degree
    above 2^20): depths / distances equal to the functors' default kernels' bit for bit, and to the oracle."""
    import numpy as np
    import oracle
    import gunrock_b200 as gb
    ro, ci = oracle.rmat_csr(18, 16, 0x5EED22)
    w = oracle.edge_weights(3, ro, ci, True)
    G = gb.graph_t.from_csr(ro, ci, w, symmetric=True)
    src = int(np.diff(ro).argmax())
    opt = gb.options_t(advance_load_balance=gb.load_balance_t.merge_path)
    d0, f0 = np.empty(G.n_vertices, np.int32), np.empty(G.n_vertices, np.float32)
    os.environ.pop("B2G_ADVANCE_VARIANT", None)
    gb.bfs(G, src, d0, options=opt)
    gb.sssp(G, src, f0, options=opt)
    variant_env(variant)
    d1, f1 = np.empty_like(d0), np.empty_like(f0)
    st = gb.bfs(G, src, d1, options=opt)
    gb.sssp(G, src, f1, options=opt)
    assert max(st.level_edges) > (1 << 20)                     # merge_path was in play
    assert np.array_equal(d0, d1) and np.array_equal(d1, oracle.bfs(ro, ci, src))
    assert np.array_equal(f0.view(np.uint32), f1.view(np.uint32))
    assert np.array_equal(f1.view(np.uint32), oracle.sssp(ro, ci, w, src).view(np.uint32))
    G.close()


DO_WORKER = r"""
import json, sys
sys.path.insert(0, %r)
import numpy as np
import oracle
import gunrock_b200 as gb
res = {}
for scale, ef, seed in ((10, 8, 3), (15, 16, 0x5EED26), (17, 8, 9)):
    ro, ci = oracle.rmat_csr(scale, ef, seed)
    G = gb.graph_t.from_csr(ro, ci, None, symmetric=True)
    deg = np.diff(ro)
    for src in (int(deg.argmax()), int(np.flatnonzero(deg > 0)[-1])):
        exp = oracle.bfs(ro, ci, src)
        for direction in (gb.advance_direction_t.optimized, gb.advance_direction_t.backward):
            d = np.empty(G.n_vertices, np.int32)
            st = gb.bfs(G, src, d, options=gb.options_t(advance_direction=direction))
            res[f"s{scale}/src{src}/dir{direction}"] = [bool(np.array_equal(d, exp)), st.level_direction]
    G.close()
print("RESULT " + json.dumps(res))
"""


@pytest.mark.parametrize("legacy", [False, True])
def test_pull_kernel_generations_give_the_same_depths(built, legacy):
    """Direction-optimised and pull-only BFS with the first-in-neighbour kernels (default) and with the
    first-generation sweep / list kernels (B2G_BFS_PULL_LEGACY=1; env read once per process: subprocess)."""
    env = dict(os.environ)
    env.pop("B2G_BFS_PULL_LEGACY", None)
    if legacy:
        env["B2G_BFS_PULL_LEGACY"] = "1"
    r = subprocess.run([sys.executable, "-c", DO_WORKER % ROOT], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert all(v[0] for v in out.values()), {k: v for k, v in out.items() if not v[0]}
    assert any(1 in v[1] for v in out.values())            # pull levels did run
