#!/usr/bin/env python3
"""How long should the speculative blocks of the row-stripe solver be?  (faldoi-ipol_b200/csrc/stripes.inc)

    python tools/stripes_policy.py trace   # CPU oracle on the bench's 3840x2160 pair -> /tmp/err_trace.txt (3-4 min)
    python tools/stripes_policy.py search  # replay block policies on that trace with a cost model

The exit test of a warp is examined once per block of launches; a block costs a checkpoint + rendezvous + host
round trip, and the block that contains the exit is partly thrown away and replayed.  `search` replays a policy
(first block length; block lengths by err / tol^2 thresholds) on the per-iteration error curve of every warp, at
1x, 2x and 4x speed (every 2nd / 4th sample: faster-converging frames), and prints the policy with the smallest
worst-case cost relative to the best fixed length.  Test tooling: uses oracle/, never the product path.
"""
import ctypes
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
TRACE = "/tmp/err_trace.txt"
LAUNCH, BLOCK, RESTORE, WARP = 30.0, 50.0, 25.0, 150.0  # microseconds (8 stripes of a 4K frame)
TOL2 = 1e-4


def trace():
    import torch
    import bench
    import pyoracle as po
    d = bench.make_pairs(1, 3840, 2160, 2000, torch.device("cpu"))
    I0, I1, u0 = (d[k][0].numpy() for k in ("I0", "I1", "u0"))
    lib = po.oracle()
    lib.fo_set_err_trace.argtypes = [ctypes.c_char_p]
    lib.fo_set_err_trace(TRACE.encode())
    print(po.o_tvl2(I0, I1, u0, warps=5)[2])
    lib.fo_set_err_trace(None)


def load(speed):
    tr = {}
    for line in open(TRACE):
        w, _, e = line.split()
        tr.setdefault(int(w), []).append(float(e))
    return [v[speed - 1::speed] + ([] if len(v) % speed == 0 else [v[-1]]) for v in tr.values()]


def cost(curves, policy, max_iters=400):
    total = 0.0
    for e in curves:
        nfull, L, last, t = (max_iters + 1) // 2, 0, 0.0, WARP
        while L < nfull:
            nb = min(policy(last), nfull - L)
            first, cnt = 2 * L, min(2 * nb, max_iters - 2 * L)
            t += BLOCK + nb * LAUNCH
            if len(e) <= first + cnt:  # the exit lies in this block
                c = len(e) - first
                if c < cnt:
                    t += RESTORE + (c + 1) // 2 * LAUNCH
                break
            last = e[first + cnt - 1]
            L += nb
        total += t
    return total


def by_ratio(start, th, ks):
    def policy(last):
        if not last > 0:
            return start
        r = last / TOL2
        return ks[0] if r > th[0] else ks[1] if r > th[1] else ks[2] if r > th[2] else ks[3]
    return policy


def search():
    speeds = {c: load(c) for c in (1, 2, 4)}
    base = {c: min(cost(v, lambda last, K=K: K) for K in (4, 8, 16)) for c, v in speeds.items()}
    best = None
    for th in itertools.product([4, 8, 16, 32], [2, 3, 4, 6], [1.2, 1.5, 2, 2.5]):
        if not th[0] > th[1] > th[2]:
            continue
        for ks in itertools.product([12, 16, 24, 32], [8, 12, 16], [4, 6, 8, 12], [2, 3, 4, 6, 8]):
            for start in (8, 16, 32):
                worst = max(cost(v, by_ratio(start, th, ks)) / base[c] for c, v in speeds.items())
                if best is None or worst < best[0]:
                    best = (worst, start, th, ks)
    print("best fixed length (us):", base)
    print("minimax policy: start %d, thresholds %s, lengths %s: %.3f of the best fixed length at worst" % (best[1], best[2], best[3], best[0]))


if __name__ == "__main__":
    {"trace": trace, "search": search}[sys.argv[1] if len(sys.argv) > 1 else "search"]()
