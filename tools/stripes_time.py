#!/usr/bin/env python3
"""Row stripes: one 3840x2160 TVL2 pair on 1 stripe and on N stripes (one per GPU), plus the steady-state time of a
two-iteration launch with the exit test out of the way (tol = 0: every warp runs all 400 iterations).

    python tools/stripes_time.py [N]        (default: every GPU of the box)
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

fb = importlib.import_module("faldoi-ipol_b200")
w, h = 3840, 2160
n = int(sys.argv[1]) if len(sys.argv) > 1 else fb.device_count()
d = bench.make_pairs(1, w, h, 2000, torch.device("cuda", 0))
host = {k: d[k][0].cpu().pin_memory() for k in ("I0", "I1", "u0")}
del d


def run(nstripes, tol=None, reps=3):
    g = fb.Stripes(w, h, list(range(nstripes)))
    params = fb.default_params(0, 400, 5)
    if tol is not None:
        params.tol = tol
    ms = []
    for r in range(reps + 1):
        g.upload_ptrs(host["I0"].data_ptr(), host["I1"].data_ptr(), host["u0"].data_ptr())
        g.run(params)
        if r:
            ms.append(g.last_run_ms)
    u, log = g.download()
    la = g.last_launches
    g.close()
    return min(ms), la, sum(log.iters[:5]), u


one = run(1)
print("1 stripe: %.2f ms, %d kernel launches, %d iterations" % one[:3])
if n > 1:
    many = run(n)
    print("%d stripes: %.2f ms (%.2fx, efficiency %.3f), %d kernel launches, bit-identical: %s"
          % (n, many[0], one[0] / many[0], one[0] / many[0] / n, many[1], bool((many[3] == one[3]).all())))
for k in sorted({1, n}):
    r = run(k, tol=0.0, reps=2)
    print("%d stripe(s), no exit: %.2f ms for %d iterations = %.2f us per two-iteration launch" % (k, r[0], r[2], r[0] * 1e3 / (r[2] / 2)))
