"""torchrun worker for the real multi-GPU test: partitioned BFS over NCCL, checked against the oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import gunrock_b200 as gb  # noqa: E402
from gunrock_b200 import multi_gpu as mg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scale, ef, seed = 16, 16, 0x5EED22
    ro, ci = oracle.rmat_csr(scale, ef, seed)
    src = int(np.diff(ro).argmax())
    exp = oracle.bfs(ro, ci, src)
    results = {}
    # (a) partition of a host CSR, (b) the device RMAT generator producing the rank's share directly
    for name, G in (("csr", mg.PartitionedGraph.from_global_csr(ro, ci, world, rank)),
                    ("rmat", mg.PartitionedGraph.rmat(scale, ef << scale, seed, world, rank))):
        for direction in (gb.advance_direction_t.forward, gb.advance_direction_t.optimized):
            for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
                eng = mg.CudaRankEngine(G, gb.options_t(advance_load_balance=lb))
                for variant, fn, kw in (("2phase", mg.bfs_rank, {}), ("async", mg.bfs_rank_async, {}),
                                        ("async-overflow", mg.bfs_rank_async, {"cap_s": 16})):
                    d, st = fn(eng, mg.TorchDistComm(), src, total_edges=len(ci), direction=direction, **kw)
                    ok = bool(np.array_equal(d.cpu().numpy(), exp[rank::world]))
                    t = torch.tensor([int(ok)], device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    results[f"{name}/{direction}/{lb}/{variant}"] = [int(t.item()), st.level_direction]
                # NCCL exchange with the level loop in C++ (bfs_nccl.cuh): own communicator, created once per handle
                mg.nccl_connect(eng, mg.TorchDistComm())
                for rep in range(2):
                    d, st = mg.bfs_rank_nccl(eng, src, len(ci), direction)
                    ok = bool(np.array_equal(d.cpu().numpy(), exp[rank::world]))
                    t = torch.tensor([int(ok)], device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    results[f"{name}/{direction}/{lb}/nccl-cxx{rep}"] = [int(t.item()), st.level_direction]
                # exchange done by the kernels over NVLink peer memory (CUDA IPC windows), twice per window
                mg.p2p_connect(eng, mg.TorchDistComm())
                for rep in range(2):
                    d, st = mg.bfs_rank_p2p(eng, src, len(ci), direction)
                    ok = bool(np.array_equal(d.cpu().numpy(), exp[rank::world]))
                    t = torch.tensor([int(ok)], device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    results[f"{name}/{direction}/{lb}/p2p{rep}"] = [int(t.item()), st.level_direction]
                dist.barrier()
        mg.p2p_disconnect(eng, mg.TorchDistComm())
        G.close()
    # partitioned PageRank over NCCL (directed graph, in-edge partition)
    dro, dci = oracle.rmat_csr(13, 8, 99, mirror=False)
    G = mg.PartitionedGraph.from_global_csr(dro, dci, world, rank, symmetric=False, by_destination=True)
    p, iters = mg.pr_rank(mg.CudaRankEngine(G), mg.TorchDistComm())
    pe, eit = oracle.pr(dro, dci, None, 0.85, 1e-6)
    mine = pe[rank::world]
    ok = iters == eit and bool(np.all(np.abs(p.cpu().numpy() - mine) <= 1e-6 * np.abs(mine)))
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    results["pagerank"] = [int(t.item()), iters]
    # the same PageRank with the iteration loop and the collectives in C++ (b2g_part_pr_nccl)
    eng = mg.CudaRankEngine(G)
    mg.nccl_connect(eng, mg.TorchDistComm())
    p, iters = mg.pr_rank_nccl(eng)
    ok = iters == eit and bool(np.all(np.abs(p.cpu().numpy() - mine) <= 1e-6 * np.abs(mine)))
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    results["pagerank/nccl-cxx"] = [int(t.item()), iters]
    G.close()
    # weighted PageRank: in-edge rows with their own weights, fp64 row sums reduced over the ranks
    dw = oracle.edge_weights(31, dro, dci, True)
    G = mg.PartitionedGraph.from_global_csr_weighted(dro, dci, dw, world, rank, symmetric=False, by_destination=True)
    pe, eit = oracle.pr(dro, dci, dw, 0.85, 1e-6)
    mine = pe[rank::world]
    eng = mg.CudaRankEngine(G)
    mg.nccl_connect(eng, mg.TorchDistComm())
    for name, run in (("pagerank-weighted", lambda: mg.pr_rank(eng, mg.TorchDistComm(), tol=0.0, max_iter=eit)),
                      ("pagerank-weighted/nccl-cxx", lambda: mg.pr_rank_nccl(eng, tol=0.0, max_iter=eit))):
        p, iters = run()
        ok = iters == eit and bool(np.all(np.abs(p.cpu().numpy() - mine) <= 1e-6 * np.abs(mine)))
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        results[name] = [int(t.item()), iters]
    G.close()
    # partitioned SSSP over NCCL
    w = oracle.edge_weights(23, ro, ci, True)
    G = mg.PartitionedGraph.from_global_csr_weighted(ro, ci, w, world, rank)
    d, iters, relaxed = mg.sssp_rank(mg.CudaRankEngine(G), mg.TorchDistComm(), src)
    es = oracle.sssp(ro, ci, w, src)[rank::world]
    ok = bool(np.array_equal(d.cpu().numpy().view(np.uint32), es.view(np.uint32)))
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    results["sssp"] = [int(t.item()), iters]
    # the same SSSP with the iteration loop and the grouped Send / Recv in C++ (b2g_part_sssp_nccl); 8: overflow path
    eng = mg.CudaRankEngine(G)
    mg.nccl_connect(eng, mg.TorchDistComm())
    for cap in (0, 8):
        d, iters, relaxed = mg.sssp_rank_nccl(eng, src, cap)
        ok = bool(np.array_equal(d.cpu().numpy().view(np.uint32), es.view(np.uint32)))
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        results[f"sssp/nccl-cxx/cap{cap}"] = [int(t.item()), iters]
    G.close()
    if rank == 0:
        print("MG_RESULT " + json.dumps(results), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
