#!/usr/bin/env python3
"""Host-side model of the WORK of two SSSP schedules on the bench graph family (design input for
DESIGN.md section 9 item 3; not product code, not a benchmark):
  * frontier Bellman-Ford, the schedule of algorithms/sssp.hxx and of b200/sssp.cuh today;
  * near/far piles (delta-stepping with one moving threshold).
Both end in the same least fixed point (checked bit-for-bit against the oracle); what differs is how many
edges are relaxed and how many bulk-synchronous iterations that takes.
    python profiles/micro/sssp_work_model.py [scale] [edge_factor]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402


def expand(ro, ci, w, dist, frontier):
    deg = ro[frontier + 1] - ro[frontier]
    src = np.repeat(frontier, deg)
    starts = np.repeat(ro[frontier], deg)
    offs = np.arange(len(src)) - np.repeat(np.cumsum(deg) - deg, deg)
    e = starts + offs
    cand = (dist[src] + w[e]).astype(np.float32)
    return ci[e], cand


def relax(dist, dst, cand):
    """min-reduce the candidates per destination, apply, return the improved vertices"""
    order = np.lexsort((cand, dst))
    d, c = dst[order], cand[order]
    first = np.ones(len(d), bool)
    first[1:] = d[1:] != d[:-1]
    d, c = d[first], c[first]
    better = c < dist[d]
    dist[d[better]] = c[better]
    return d[better]


def bellman_ford(ro, ci, w, source):
    dist = np.full(len(ro) - 1, np.finfo(np.float32).max, np.float32)
    dist[source] = 0
    frontier = np.array([source])
    relaxed = iters = 0
    while len(frontier):
        dst, cand = expand(ro, ci, w, dist, frontier)
        relaxed += len(dst)
        frontier = relax(dist, dst, cand)
        iters += 1
    return dist, relaxed, iters


def near_far(ro, ci, w, source, delta):
    dist = np.full(len(ro) - 1, np.finfo(np.float32).max, np.float32)
    dist[source] = 0
    near, far = np.array([source]), np.zeros(0, np.int64)
    threshold = np.float32(delta)
    relaxed = iters = 0
    while len(near) or len(far):
        while len(near):
            dst, cand = expand(ro, ci, w, dist, near)
            relaxed += len(dst)
            improved = relax(dist, dst, cand)
            iters += 1
            is_near = dist[improved] < threshold
            near = improved[is_near]
            far = np.concatenate([far, improved[~is_near]])
        if len(far):
            far = np.unique(far)
            threshold = np.float32(max(threshold + np.float32(delta), 0))
            # a far entry is stale if the vertex was settled below the old threshold meanwhile: it was
            # expanded then with a distance it still has
            lo = threshold - np.float32(delta)
            far = far[dist[far] >= lo]
            take = dist[far] < threshold
            near, far = far[take], far[~take]
            if not len(near) and len(far):      # jump over empty buckets
                threshold = np.float32(dist[far].min()) + np.float32(delta)
                take = dist[far] < threshold
                near, far = far[take], far[~take]
    return dist, relaxed, iters


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    ef = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    ro, ci = oracle.rmat_csr(scale, ef, 0x5EED24)
    ro64 = ro.astype(np.int64)
    source = int(np.diff(ro).argmax())
    for mode, name in ((False, "integer weights 1..63"), (True, "weights 1 + 63 u01")):
        w = oracle.edge_weights(0x5EED25, ro, ci, mode)
        exp = oracle.sssp(ro, ci, w, source)
        d, relaxed, iters = bellman_ford(ro64, ci.astype(np.int64), w, source)
        assert np.array_equal(d.view(np.uint32), exp.view(np.uint32))
        reach = int((np.diff(ro)[exp < np.finfo(np.float32).max]).sum())
        print(f"RMAT-{scale} ef{ef}, {name}: |E| = {len(ci)}, edges of reached vertices = {reach}")
        print(f"  frontier Bellman-Ford : relaxed {relaxed:>12d} ({relaxed / reach:5.2f} x)  iterations {iters}")
        for delta in (4, 8, 16, 32, 64):
            d, relaxed, iters = near_far(ro64, ci.astype(np.int64), w, source, delta)
            assert np.array_equal(d.view(np.uint32), exp.view(np.uint32)), delta
            print(f"  near/far, delta = {delta:<3d}: relaxed {relaxed:>12d} ({relaxed / reach:5.2f} x)  iterations {iters}")


if __name__ == "__main__":
    main()
