"""Files in, files out: `global_faldoi -seq` on N synthetic 1024x436 pairs (binary PPM frames + .flo init flows
written to a scratch directory), TVL2, default parameters.  Reports pairs/s through the CLI -- start-up excluded by
differencing against a short run -- for the GPUs given:

    python tools/seq_time.py [N=64] [devices=0 | all | 0,1,..] [batch=16] [method=0]

What the pipeline does per pair: read + decode 3 frames and a flow (host thread pool), fill a slot of a batched
solver handle through pinned staging, solve, download, write the .flo on a background thread.
"""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
BIN = os.path.join(ROOT, "faldoi-ipol_b200", "bin", "global_faldoi")


def main():
    import torch
    from bench import make_pairs
    import pyoracle as po  # test infrastructure: only its .flo / PPM writers are used here
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    devices = sys.argv[2] if len(sys.argv) > 2 else "0"
    batch = sys.argv[3] if len(sys.argv) > 3 else "16"
    method = sys.argv[4] if len(sys.argv) > 4 else "0"
    w, h = 1024, 436
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=base) as t:
        d = make_pairs(n, w, h, 1000, "cuda:0" if torch.cuda.is_available() else "cpu")
        rgb = d["rgb"].clamp(0, 255).byte().cpu().numpy()
        u0 = d["u0"].cpu().numpy()
        lines = []
        for k in range(n):
            # frame k and its two "neighbours": distinct files per pair, like a real sequence
            names = []
            for j, sh in enumerate((0, 1, -1)):
                p = os.path.join(t, "p%03d_f%d.ppm" % (k, j))
                po.write_ppm(p, np.ascontiguousarray(np.roll(rgb[k], sh, 2).transpose(1, 2, 0)))
                names.append(p)
            ims = os.path.join(t, "ims%03d.txt" % k)
            open(ims, "w").write("\n".join(names) + "\n")
            flo = os.path.join(t, "in%03d.flo" % k)
            po.write_flo(flo, u0[k])
            lines.append("%s %s %s\n" % (ims, flo, os.path.join(t, "out%03d.flo" % k)))
        res = {}
        for jobs in (min(4, n), n):
            open(os.path.join(t, "jobs.txt"), "w").write("".join(lines[:jobs]))
            t0 = time.perf_counter()
            r = subprocess.run([BIN, "-seq", os.path.join(t, "jobs.txt"), "-m", method, "-w", "5", "-devices", devices, "-batch", batch, "-seq_stats", "1"],
                               capture_output=True, text=True)
            res[jobs] = time.perf_counter() - t0
            assert r.returncode == 0 and ("sequence: %d pairs done" % jobs) in r.stderr, r.stderr[-2000:]
            print("%d pairs on devices %s, batch %s: %.2f s wall" % (jobs, devices, batch, res[jobs]), flush=True)
            print("".join(m + "\n" for m in re.findall(r"(?:device \d+: waiting|host: dispatcher|pipeline: )[^\n]*", r.stderr)), end="")
        if n > 4:
            dt = res[n] - res[min(4, n)]
            print("marginal: %.4f s per pair = %.1f pairs/s (start-up + first 4 pairs: %.2f s)" % (dt / (n - 4), (n - 4) / dt, res[4]))
        print("whole run: %.1f pairs/s" % (n / res[n]))


if __name__ == "__main__":
    main()
