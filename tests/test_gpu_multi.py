"""GPU tests of the partitioned (multi-GPU) BFS.
* several simulated ranks on ONE GPU (the per-rank CUDA steps are the real ones; the exchange is a
  tensor copy) -- runs wherever a single B200 is available;
* a real NCCL run under torchrun when >= 2 GPUs are visible."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb(built):
    import gunrock_b200 as gb
    if gb.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device")
    return gb


@pytest.mark.parametrize("P", [1, 2, 3, 4, 8])
def test_simulated_ranks_on_one_gpu(gb, P):
    from gunrock_b200 import multi_gpu as mg
    for scale, ef, seed, mirror in ((11, 16, 5, True), (14, 8, 0x5EED22, True)):
        ro, ci = oracle.rmat_csr(scale, ef, seed, mirror=mirror)
        deg = np.diff(ro)
        graphs = [mg.PartitionedGraph.from_global_csr(ro, ci, P, r) for r in range(P)]
        for src in (int(deg.argmax()),) if P > 2 else (int(deg.argmax()), int(np.flatnonzero(deg > 0)[-1])):
            exp = oracle.bfs(ro, ci, src)
            for direction in (gb.advance_direction_t.forward, gb.advance_direction_t.optimized,
                              gb.advance_direction_t.backward):
                for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
                    engines = [mg.CudaRankEngine(g, gb.options_t(advance_load_balance=lb, hub_threshold=256))
                               for g in graphs]
                    dists, st = mg.bfs_lockstep(engines, src, total_edges=len(ci), direction=direction)
                    got = mg.gather_distances([d.cpu().numpy() for d in dists], len(ro) - 1)
                    assert np.array_equal(got, exp), (P, scale, src, direction, lb)
                    if direction == gb.advance_direction_t.forward:
                        # push inspects every out-edge of every reached vertex exactly once
                        assert st.edges_touched == int(deg[exp < 2**31 - 1].sum())
        for g in graphs:
            g.close()


@pytest.fixture(scope="module")
def p2p_sim_results(gb):
    """tests/p2p_sim_worker.py in a subprocess with eager module loading (see its docstring)."""
    env = dict(os.environ, CUDA_MODULE_LOADING="EAGER", B2G_P2P_TIMEOUT_MS="5000")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "p2p_sim_worker.py")], capture_output=True,
                       text=True, timeout=1500, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("P2P_SIM_RESULT ")][-1]
    return json.loads(line[len("P2P_SIM_RESULT "):])


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_peer_memory_exchange_simulated_ranks(p2p_sim_results, P):
    """bfs_p2p.cuh: the kernels write forwarded ids / frontier words / statistics straight into the
    peers' windows and synchronise with epoch flags -- here the peers are P handles on one GPU (one
    host thread and one stream per rank); the kernels and the protocol are the multi-GPU ones.
    Depths bit-exact against the oracle for 2 graphs x 2 load balancers x (2 | 1) sources x 3 directions."""
    mine = {k: v for k, v in p2p_sim_results.items() if k.startswith(f"P{P}/")}
    assert len(mine) == 18
    assert all(v[0] == 1 for v in mine.values()), {k: v for k, v in mine.items() if v[0] != 1}
    # the pull path (sweep writing into the peers' frontier maps) really ran
    assert any(1 in v[1] for k, v in mine.items() if k.endswith("dir2"))


def test_partitioned_rmat_generator_matches_global(gb):
    from gunrock_b200 import multi_gpu as mg
    scale, ef, seed, P = 13, 8, 99, 4
    ro, ci = oracle.rmat_csr(scale, ef, seed)
    graphs = [mg.PartitionedGraph.rmat(scale, ef << scale, seed, P, r) for r in range(P)]
    src = int(np.diff(ro).argmax())
    engines = [mg.CudaRankEngine(g) for g in graphs]
    dists, st = mg.bfs_lockstep(engines, src, total_edges=len(ci))
    got = mg.gather_distances([d.cpu().numpy() for d in dists], len(ro) - 1)
    assert np.array_equal(got, oracle.bfs(ro, ci, src))
    assert sum(g.n_local for g in graphs) == len(ro) - 1


def test_nccl_two_or_more_gpus(gb):
    n = gb.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MG_RESULT ")][-1]
    res = json.loads(line[len("MG_RESULT "):])
    assert len(res) == 63 and all(v[0] == 1 for v in res.values()), res


def test_cxx_nccl_level_loop_single_rank(gb):
    """world_size 1: the C++ level loop with the NCCL exchange (b2g_part_bfs_nccl) -- communicator created by
    ncclCommInitRank inside the library from an id made by the library, all collectives skipped at one rank, the
    same kernels / statistics / pinned-memory polling as at N ranks.  (N > 1: tests/mg_worker.py under torchrun.)"""
    from gunrock_b200 import multi_gpu as mg
    ro, ci = oracle.rmat_csr(15, 16, 77)
    G = mg.PartitionedGraph.from_global_csr(ro, ci, 1, 0)
    deg = np.diff(ro)
    for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
        eng = mg.CudaRankEngine(G, gb.options_t(advance_load_balance=lb))
        eng.nccl_init(mg.nccl_unique_id())
        for src in (int(deg.argmax()), int(np.flatnonzero(deg > 0)[-1])):
            exp = oracle.bfs(ro, ci, src)
            for direction in (gb.advance_direction_t.forward, gb.advance_direction_t.optimized,
                              gb.advance_direction_t.backward):
                d, st = mg.bfs_rank_nccl(eng, src, len(ci), direction)
                assert np.array_equal(d.cpu().numpy(), exp), (lb, src, direction)
                assert st.levels == int(exp[exp < 2**31 - 1].max()) + 1
    G.close()


def test_cxx_nccl_sssp_and_pagerank_single_rank(gb):
    """world_size 1: the C++ iteration loops of the partitioned SSSP and PageRank (b2g_part_sssp_nccl,
    b2g_part_pr_nccl): same kernels, statistics records and stopping rules as at N ranks, collectives skipped.
    (N > 1: tests/mg_worker.py under torchrun.)"""
    from gunrock_b200 import multi_gpu as mg
    ro, ci = oracle.rmat_csr(13, 8, 606)
    w = oracle.edge_weights(17, ro, ci, True)
    deg = np.diff(ro)
    G = mg.PartitionedGraph.from_global_csr_weighted(ro, ci, w, 1, 0)
    for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
        eng = mg.CudaRankEngine(G, gb.options_t(advance_load_balance=lb, hub_threshold=256))
        eng.nccl_init(mg.nccl_unique_id())
        for src in (int(deg.argmax()), int(np.flatnonzero(deg > 0)[-1])):
            exp = oracle.sssp(ro, ci, w, src)
            for cap in (0, 8):    # 8: a message row overflows -> the run restarts with longer rows
                d, it, relaxed = mg.sssp_rank_nccl(eng, src, cap)
                assert np.array_equal(d.cpu().numpy().view(np.uint32), exp.view(np.uint32)), (lb, src, cap)
                assert it > 0 and relaxed > 0
    G.close()
    # PageRank: unweighted (in-edge rows of a directed graph) and weighted
    ro, ci = oracle.rmat_csr(12, 8, 4321, mirror=False)
    w = oracle.edge_weights(5, ro, ci, True)
    for weights in (None, w):
        if weights is None:
            G = mg.PartitionedGraph.from_global_csr(ro, ci, 1, 0, symmetric=False, by_destination=True)
        else:
            G = mg.PartitionedGraph.from_global_csr_weighted(ro, ci, w, 1, 0, symmetric=False, by_destination=True)
        eng = mg.CudaRankEngine(G)
        eng.nccl_init(mg.nccl_unique_id())
        exp, exp_iters = oracle.pr(ro, ci, weights, 0.85, 1e-6)
        p, it = mg.pr_rank_nccl(eng)
        assert abs(it - exp_iters) <= (0 if weights is None else 1)
        p, it = mg.pr_rank_nccl(eng, tol=0.0, max_iter=exp_iters)
        assert it == exp_iters
        rel = np.abs(p.cpu().numpy() - exp) / np.maximum(np.abs(exp), np.finfo(np.float32).tiny)
        assert rel.max() <= 1e-6, (weights is not None, rel.max())
        G.close()


@pytest.mark.parametrize("P", [1, 2, 4])
def test_partitioned_weighted_pagerank_simulated_ranks(gb, P):
    """Weighted PageRank over a partition by destination: every in-edge carries its own weight, the row sums of
    the weights are reduced over the ranks in fp64 (b2g_part_pr_outweights / b2g_part_pr_begin_weighted)."""
    from gunrock_b200 import multi_gpu as mg
    for mirror in (True, False):
        ro, ci = oracle.rmat_csr(12, 8, 99, mirror=mirror)
        w = oracle.edge_weights(23, ro, ci, True)      # w(u->v) != w(v->u) in general, also when mirrored
        graphs = [mg.PartitionedGraph.from_global_csr_weighted(ro, ci, w, P, r, symmetric=mirror, by_destination=True)
                  for r in range(P)]
        exp, exp_iters = oracle.pr(ro, ci, w, 0.85, 1e-6)
        ps, iters = mg.pr_lockstep([mg.CudaRankEngine(g) for g in graphs])
        assert abs(iters - exp_iters) <= 1
        ps, iters = mg.pr_lockstep([mg.CudaRankEngine(g) for g in graphs], tol=0.0, max_iter=exp_iters)
        got = mg.gather_distances([p.cpu().numpy() for p in ps], len(ro) - 1)
        rel = np.abs(got - exp) / np.maximum(np.abs(exp), np.finfo(np.float32).tiny)
        assert iters == exp_iters and rel.max() <= 1e-6, (P, mirror, rel.max())
        for g in graphs:
            g.close()


def test_async_driver_single_rank(gb):
    """world_size 1 under NCCL: the sync-free level loop (one host sync per level) on one GPU."""
    import torch
    import torch.distributed as dist
    from gunrock_b200 import multi_gpu as mg
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29622")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        ro, ci = oracle.rmat_csr(15, 16, 77)
        G = mg.PartitionedGraph.from_global_csr(ro, ci, 1, 0)
        src = int(np.diff(ro).argmax())
        exp = oracle.bfs(ro, ci, src)
        for direction in (gb.advance_direction_t.forward, gb.advance_direction_t.optimized):
            eng = mg.CudaRankEngine(G)
            d, st = mg.bfs_rank_async(eng, mg.TorchDistComm(), src, len(ci), direction=direction)
            assert np.array_equal(d.cpu().numpy(), exp)
            d, st = mg.bfs_rank(eng, mg.TorchDistComm(), src, len(ci), direction=direction)
            assert np.array_equal(d.cpu().numpy(), exp)
        G.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("mirror", [True, False])
def test_partitioned_pagerank_simulated_ranks(gb, P, mirror):
    """Partitioned pull PageRank (owner of the destination pulls; c all-gathered) vs the oracle."""
    from gunrock_b200 import multi_gpu as mg
    ro, ci = oracle.rmat_csr(12, 8, 4321, mirror=mirror)
    graphs = [mg.PartitionedGraph.from_global_csr(ro, ci, P, r, symmetric=mirror, by_destination=True)
              for r in range(P)]
    engines = [mg.CudaRankEngine(g) for g in graphs]
    ps, iters = mg.pr_lockstep(engines)
    got = mg.gather_distances([p.cpu().numpy() for p in ps], len(ro) - 1)
    exp, exp_iters = oracle.pr(ro, ci, None, 0.85, 1e-6)
    assert iters == exp_iters
    rel = np.abs(got - exp) / np.maximum(np.abs(exp), np.finfo(np.float32).tiny)
    assert rel.max() <= 1e-6, rel.max()
    for g in graphs:
        g.close()
    # the device generator's in-edge partition of a DIRECTED graph reproduces the same ranks
    if not mirror:
        graphs = [mg.PartitionedGraph.rmat(12, 8 << 12, 4321, P, r, mirror=False, by_destination=True)
                  for r in range(P)]
        ps, iters = mg.pr_lockstep([mg.CudaRankEngine(g) for g in graphs])
        got2 = mg.gather_distances([p.cpu().numpy() for p in ps], len(ro) - 1)
        assert iters == exp_iters and np.array_equal(got2, got)
        for g in graphs:
            g.close()


@pytest.mark.parametrize("P", [1, 2, 4])
def test_partitioned_sssp_simulated_ranks(gb, P):
    """Partitioned SSSP: (vertex, distance) pairs pushed to the owner; fp32 distances bit-exact."""
    from gunrock_b200 import multi_gpu as mg
    ro, ci = oracle.rmat_csr(13, 8, 606)
    w = oracle.edge_weights(17, ro, ci, True)
    graphs = [mg.PartitionedGraph.from_global_csr_weighted(ro, ci, w, P, r) for r in range(P)]
    deg = np.diff(ro)
    for src in (int(deg.argmax()), int(np.flatnonzero(deg > 0)[-1])):
        exp = oracle.sssp(ro, ci, w, src)
        for cap in (0, 8):        # 8: forces the overflow -> restart path
            engines = [mg.CudaRankEngine(g, gb.options_t(hub_threshold=256)) for g in graphs]
            ds, iters = mg.sssp_lockstep(engines, src, cap_s=cap)
            got = mg.gather_distances([d.cpu().numpy() for d in ds], len(ro) - 1)
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (P, src, cap)
    for g in graphs:
        g.close()
