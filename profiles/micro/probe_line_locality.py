#!/usr/bin/env python
"""How many 128-byte lines of the visited bitmap does ONE warp-wide probe touch?  (design input, host only)

The push-BFS kernels probe one visited bit per edge; a warp's 32 probes cost one L1TEX wavefront per DISTINCT
128-byte line (1024 vertices).  This script builds the bench graph family on the host (oracle generator), takes the
level-1 frontier of the bench source and counts distinct lines per 32-edge chunk for
  A) the kernels' mapping: 32 consecutive edges of one (sorted) row per warp instruction,
  B) a thread-per-row mapping inside power-of-two degree classes.
Result (profiles/r1_probe_line_locality.txt): rows are sorted and RMAT ids are skewed towards 0, so A touches
only 8-12 lines per chunk at scales 18-21 (~14 extrapolated to scale 22), not 32 -- the L1TEX wavefront rate of
the level-1 launch is then ~0.45 x 84.7 M / 378 us = 100 G/s, about a third of 148 SMs x 1.97 GHz, which is what ncu
reports (l1tex throughput 40 %).  B is worse than A.  Usage: python profiles/micro/probe_line_locality.py [scale]"""
import sys, time
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, oracle
scale=int(sys.argv[1]) if len(sys.argv)>1 else 20
t=time.time()
ro,ci=oracle.rmat_csr(scale,16,0x5EED22)
V=len(ro)-1; deg=np.diff(ro)
print('graph', V, len(ci), 'build s', round(time.time()-t,1))
src=int(deg.argmax())
front=ci[ro[src]:ro[src+1]]
print('L1 frontier rows', len(front), 'edges', int(deg[front].sum()), 'avg deg', deg[front].mean())
LINE=10  # 1024 vertices per 128-byte bitmap line
# A) row-major chunks of 32 consecutive edges
rng=np.random.default_rng(0)
sample=rng.choice(front, min(len(front),20000), replace=False)
linesA=0; chunksA=0; edgesA=0
for v in sample:
    row=ci[ro[v]:ro[v+1]]>>LINE
    n=len(row)
    for c in range(0,n,32):
        seg=row[c:c+32]
        linesA+=len(np.unique(seg)); chunksA+=1; edgesA+=len(seg)
print(f'A row-major: {linesA/edgesA:.3f} probe wavefronts per edge ({linesA/chunksA:.1f} lines per 32-edge chunk, fill {edgesA/chunksA:.1f})')
# B) thread-per-row inside degree classes
linesB=0; edgesB=0; stepsB=0; col_wf=0
d=deg[sample]
for j in range(0,20):
    cls=sample[(d>=2**j)&(d<2**(j+1))]
    if len(cls)==0: continue
    lc=0; ec=0
    for w in range(0,len(cls),32):
        rows=[ci[ro[v]:ro[v+1]]>>LINE for v in cls[w:w+32]]
        m=max(len(r) for r in rows)
        for k in range(m):
            ids=[r[k] for r in rows if k<len(r)]
            lc+=len(set(ids)); ec+=len(ids); stepsB+=1
    linesB+=lc; edgesB+=ec
    print(f'  class 2^{j}: rows {len(cls)} edges {ec} probe wf/edge {lc/max(ec,1):.3f}')
print(f'B thread-per-row (degree classes): {linesB/edgesB:.3f} probe wavefronts per edge; lane utilisation {edgesB/(stepsB*32):.2f}')
