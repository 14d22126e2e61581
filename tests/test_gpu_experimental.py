"""Experimental paths that are OFF by default and were written without GPU time left in round 1.
They only run when B2G_RUN_EXPERIMENTAL=1 is set (first thing to do on a GPU box next round):

    B2G_RUN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B2G_RUN_EXPERIMENTAL") != "1", reason="experimental, opt-in")]

WORKER = r"""
import json, os, sys
sys.path.insert(0, %r)
import numpy as np
import oracle
import gunrock_b200 as gb
res = {}
for scale, ef, seed, mode in ((10, 8, 3, False), (14, 16, 0x5EED24, False), (15, 8, 9, True)):
    ro, ci = oracle.rmat_csr(scale, ef, seed)
    w = oracle.edge_weights(seed + 1, ro, ci, mode)
    G = gb.graph_t.from_csr(ro, ci, w, symmetric=True)
    deg = np.diff(ro)
    for src in (int(deg.argmax()), int(np.flatnonzero(deg > 0)[-1])):
        exp = oracle.sssp(ro, ci, w, src)
        for lb in (gb.load_balance_t.block_mapped, gb.load_balance_t.merge_path):
            d = np.empty(G.n_vertices, np.float32)
            st = gb.sssp(G, src, d, options=gb.options_t(advance_load_balance=lb, hub_threshold=256))
            res[f"s{scale}/src{src}/lb{lb}"] = [bool(np.array_equal(d.view(np.uint32), exp.view(np.uint32))),
                                               st.iterations, int(st.edges_touched)]
    G.close()
print("RESULT " + json.dumps(res))
"""


@pytest.mark.parametrize("delta", ["4", "8", "16.5", "1000"])
def test_sssp_near_far_schedule_is_bit_exact(built, delta):
    """sssp_run_near_far (B2G_SSSP_DELTA): same least fixed point as the oracle, fewer relaxations than the
    default schedule for small delta (the env var is read once per process, hence the subprocess)."""
    out = {}
    for d in (delta, ""):
        env = dict(os.environ)
        env.pop("B2G_SSSP_DELTA", None)
        if d:
            env["B2G_SSSP_DELTA"] = d
        r = subprocess.run([sys.executable, "-c", WORKER % ROOT], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[d] = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert all(v[0] for v in out[delta].values()), {k: v for k, v in out[delta].items() if not v[0]}
    assert all(v[0] for v in out[""].values())
    if float(delta) <= 8:
        big = [k for k in out[delta] if k.startswith("s14") or k.startswith("s15")]
        assert sum(out[delta][k][2] for k in big) < sum(out[""][k][2] for k in big)     # less work
