import os, sys
os.environ["B2G_TRACE"] = "1"
os.environ.setdefault("B2G_P2P_TIMEOUT_MS", "5000")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gunrock_b200 as gb
from gunrock_b200 import multi_gpu as mg
P, scale = int(sys.argv[1]), int(sys.argv[2])
graphs = [mg.PartitionedGraph.rmat(scale, 16 << scale, 0x5EED22, P, r, mirror=True) for r in range(P)]
best = max((g.max_degree_vertex()[1], g.max_degree_vertex()[0]) for g in graphs)
src = best[1]
total = sum(g.n_local_edges for g in graphs)
engines = [mg.CudaRankEngine(g, gb.options_t(advance_load_balance=gb.load_balance_t.block_mapped)) for g in graphs]
mg.p2p_connect_simulated(engines)
for rep in range(3):
    dists, st = mg.bfs_threads_p2p(engines, src, total, gb.advance_direction_t.optimized)
print("levels", st.level_direction, st.level_frontier, flush=True)
