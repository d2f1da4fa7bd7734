"""Device Lab conversion / full raw path vs the reference (full-size Sintel pair in oracle/_ref/data)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
fb = importlib.import_module("faldoi-ipol_b200")
import pyoracle as po
D = os.path.join(ROOT, "oracle", "_ref", "data", "clean_easy")
fr = [po.read_image_planar(os.path.join(D, "frame_%04d.png" % k)) for k in (1, 2, 3)]
u0 = po.read_flo(os.path.join(D, "rg.flo"))
h, w = u0.shape[1:]
s = fb.Solver(w, h, 7, 1)
s.upload_raw(0, fr[1], fr[2], fr[0], u0)
I0n, I1n, _, lab = s.download_frames(0)
hl = po.o_image_to_lab(fr[1])
d = lab != hl
print("Lab values that differ from the host's: %d of %d (per channel %s), max |d| %.3g" % (d.sum(), d.size, d.reshape(3, -1).sum(1), np.abs(lab - hl).max()))
s.close()
ref = po.read_flo(os.path.join(D, "var_m7.flo"))
u, _, its, _ = fb.global_solve_raw(7, fr[1], fr[2], fr[0], u0, warps=5, glb_iters=400)
dd = np.abs(u - ref)
print("m7 raw path vs reference flow: equal %s, max %.3g, differing values %d" % (np.array_equal(u, ref), dd.max(), (dd > 0).sum()))
