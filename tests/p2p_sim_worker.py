"""Simulated ranks of the peer-memory BFS on ONE GPU (run by tests/test_gpu_multi.py in a subprocess).

Several ranks share one process and one device here, so their barrier kernels spin while the other
ranks' kernels must be able to start.  CUDA's lazy module loading synchronises the context when a
kernel is launched for the first time, which would dead-lock against a spinning kernel: the parent
starts this worker with CUDA_MODULE_LOADING=EAGER.  (With one process per GPU -- the real deployment --
every rank has its own context and nothing of this applies.)  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import oracle  # noqa: E402  (the checker: tests only)
import gunrock_b200 as gb  # noqa: E402
from gunrock_b200 import multi_gpu as mg  # noqa: E402


def main():
    results = {}
    for P in (1, 2, 3, 8):
        for scale, ef, seed in ((11, 16, 5), (14, 8, 0x5EED22)):
            ro, ci = oracle.rmat_csr(scale, ef, seed, mirror=True)
            deg = np.diff(ro)
            graphs = [mg.PartitionedGraph.from_global_csr(ro, ci, P, r) for r in range(P)]
            for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
                engines = [mg.CudaRankEngine(g, gb.options_t(advance_load_balance=lb, hub_threshold=256))
                           for g in graphs]
                mg.p2p_connect_simulated(engines)          # idempotent: the windows are reused
                srcs = (int(deg.argmax()), int(np.flatnonzero(deg > 0)[-1])) if scale < 14 else (int(deg.argmax()),)
                for src in srcs:
                    exp = oracle.bfs(ro, ci, src)
                    for direction in (gb.advance_direction_t.forward, gb.advance_direction_t.optimized,
                                      gb.advance_direction_t.backward):
                        dists, st = mg.bfs_threads_p2p(engines, src, len(ci), direction)
                        got = mg.gather_distances([d.cpu().numpy() for d in dists], len(ro) - 1)
                        ok = bool(np.array_equal(got, exp))
                        if direction == gb.advance_direction_t.forward:   # every out-edge of every reached vertex once
                            ok = ok and st.edges_touched == int(deg[exp < 2**31 - 1].sum())
                        results[f"P{P}/s{scale}/lb{lb}/src{src}/dir{direction}"] = [int(ok), st.level_direction]
            for g in graphs:
                g.close()
    print("P2P_SIM_RESULT " + json.dumps(results), flush=True)


if __name__ == "__main__":
    main()
