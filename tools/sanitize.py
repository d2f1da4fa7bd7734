"""Small solves of every kernel family, for `compute-sanitizer` (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize.py

Ragged multi-tile frames, so that tile aprons, TMA zero fill and the frame-border branches are all exercised; the
striped solve runs three stripes on device 0 (peer stores, boundary signals, speculative blocks with rollback)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_pair  # noqa: E402

fb = importlib.import_module("faldoi-ipol_b200")


def main():
    w, h = 261, 43
    I0, I1, Im1, u0, rgb = synthetic_pair(w, h, seed=5)
    import ctypes as C
    hostlib = C.CDLL(os.path.join(ROOT, "faldoi-ipol_b200", "libfaldoi_host.so"))
    lab = np.empty_like(rgb)
    hostlib.faldoi_host_image_to_lab(rgb.ctypes.data_as(C.c_void_p), w, h, lab.ctypes.data_as(C.c_void_p))
    chi = np.zeros((h, w), np.float32)
    for method, iters in ((0, 40), (4, 12), (2, 6), (6, 6), (8, 3)):
        p = fb.default_params(method, glb_iters=iters, warps=2)
        p.max_iters = iters
        u, _, its, _ = fb.global_solve(method, I0, I1, u0, Im1=Im1 if method == 8 else None, lab=lab if method in (2, 6) else None,
                                       chi=chi if method == 8 else None, params=p)
        assert np.isfinite(u).all()
        print("method", method, "iterations", its, flush=True)
    p = fb.default_params(0, warps=2)
    whole, _, its, _ = fb.global_solve(0, I0, I1, u0, params=p)
    g = fb.Stripes(w, h, [0, 0, 0])
    g.upload(I0, I1, u0)
    g.run(p)
    u, log = g.download()
    assert np.array_equal(u, whole) and list(log.iters[:2]) == its
    g.close()
    print("stripes ok", its, flush=True)


if __name__ == "__main__":
    main()
