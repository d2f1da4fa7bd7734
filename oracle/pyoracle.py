"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle and (when built)
the compiled reference.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.

  oracle()  -> oracle/_build/libfaldoi_oracle.so  (our C restatement; `make -C oracle oracle`)
  ref()     -> oracle/_ref/libfaldoi_ref.so       (UNMODIFIED reference, `make -C oracle ref`,
                                                   only buildable where /root/reference exists)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libfaldoi_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfaldoi_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "global_faldoi")
MAX_WARPS = 64

fp = C.POINTER(C.c_float)


class FoLog(C.Structure):
    _fields_ = [("iters", C.c_int * MAX_WARPS), ("err", C.c_float * MAX_WARPS)]


class FoParams(C.Structure):
    _fields_ = [(k, C.c_float) for k in
                ("lambda_", "theta", "tau", "beta", "alpha", "tau_u", "tau_eta", "tau_chi", "tol", "mu")]


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(fp)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        _oracle = C.CDLL(ORACLE_SO)
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
    return _ref


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def default_params():
    p = FoParams()
    oracle().fo_default_params(C.byref(p))
    return p


# ---------------------------------------------------------------- oracle side
def o_centered_gradient(f):
    h, w = f.shape
    dx, dy = np.empty_like(f), np.empty_like(f)
    oracle().fo_centered_gradient(_p(f), _p(dx), _p(dy), w, h)
    return dx, dy


def o_forward_gradient(f):
    h, w = f.shape
    dx, dy = np.empty_like(f), np.empty_like(f)
    oracle().fo_forward_gradient(_p(f), _p(dx), _p(dy), w, h)
    return dx, dy


def o_divergence(a, b):
    h, w = a.shape
    d = np.empty_like(a)
    oracle().fo_divergence(_p(a), _p(b), _p(d), w, h)
    return d


def o_bicubic_warp(img, u, v, border_out):
    h, w = img.shape
    out = np.empty_like(img)
    oracle().fo_bicubic_warp(_p(img), _p(u), _p(v), _p(out), w, h, int(border_out))
    return out


def o_preprocess(i0, i1, im1):
    """planar (pd,h,w) float32 0..255 -> normalised+smoothed gray triple"""
    pd, h, w = i0.shape
    outs = [np.empty((h, w), np.float32) for _ in range(3)]
    oracle().fo_preprocess(_p(i0), _p(i1), _p(im1), pd, w, h, *[_p(o) for o in outs])
    return outs


def o_image_to_lab(rgb):
    pd, h, w = rgb.shape
    assert pd == 3
    lab = np.empty_like(rgb)
    oracle().fo_image_to_lab(_p(rgb), w * h, _p(lab))
    return lab


def o_global_solve(method, I0, I1, Im1, lab, u, chi=None, params=None, warps=5, glb_iters=400):
    """u: (2,h,w) init flow. Returns (u_out, chi_out, iters, errs)."""
    h, w = I0.shape
    u = f32(u).copy()
    chi_a = f32(chi).copy() if chi is not None else None
    p = params or default_params()
    log = FoLog()
    rc = oracle().fo_global_solve(
        int(method), _p(I0), _p(I1), _p(Im1) if Im1 is not None else None,
        _p(lab) if lab is not None else None, _p(u), _p(chi_a) if chi_a is not None else None,
        C.byref(p), w, h, int(warps), int(glb_iters), C.byref(log))
    if rc != 0:
        raise RuntimeError("fo_global_solve failed")
    return u, chi_a, list(log.iters[:warps]), list(log.err[:warps])


def o_tvl2(I0, I1, u, xi=None, lam=40.0, theta=0.3, tau=0.125, tol=0.01, warps=5, max_iter=400):
    h, w = I0.shape
    u = f32(u).copy()
    xi = np.zeros((4, h, w), np.float32) if xi is None else f32(xi).copy()
    log = FoLog()
    fl = C.c_float
    oracle().fo_tvl2(_p(I0), _p(I1), _p(u[0]), _p(u[1]), _p(xi[0]), _p(xi[1]), _p(xi[2]), _p(xi[3]),
                     fl(lam), fl(theta), fl(tau), fl(tol), w, h, int(warps), int(max_iter), C.byref(log))
    return u, xi, list(log.iters[:warps]), list(log.err[:warps])


# ------------------------------------------------------------- reference side
def r_global_solve(method, I0, I1, Im1, lab, u, chi=None, warps=5, glb_iters=400, params_file=""):
    """Calls the UNMODIFIED reference solver for `method` with main()'s constants
    (src/global_faldoi.cpp:2132-2167).  Reference solvers always run 400 iterations
    for methods 0-7 (compile-time MAX_ITERATIONS_GLOBAL)."""
    L = ref()
    h, w = I0.shape
    u = f32(u).copy()
    u1, u2 = u[0], u[1]
    fl = C.c_float
    I1c = I1.copy()
    chi_a = None
    if method in (0, 1, 4, 5):
        xi = np.zeros((4, h, w), np.float32)
        if method <= 1:
            L.ref_tvl2OF(_p(I0), _p(I1c), _p(u1), _p(u2), _p(xi[0]), _p(xi[1]), _p(xi[2]), _p(xi[3]),
                         fl(40.0), fl(0.3), fl(0.125), fl(0.01), w, h, int(warps), 0)
        else:
            L.ref_tvcsad_PD(_p(I0), _p(I1c), _p(xi[0]), _p(xi[1]), _p(xi[2]), _p(xi[3]),
                            fl(0.85), fl(0.3), fl(0.125), fl(0.01), w, h, int(warps), 0, _p(u1), _p(u2))
    elif method in (2, 3):
        labc = lab.copy()
        L.ref_nltvl1_PD(_p(I0), _p(I1c), _p(labc), 3, fl(2.0), fl(0.3), fl(0.1), w, h, int(warps), 0,
                        _p(u1), _p(u2))
    elif method in (6, 7):
        labc = lab.copy()
        L.ref_nltvcsad_PD(_p(I0), _p(I1c), _p(labc), 3, fl(0.85), fl(0.3), fl(0.1), w, h, int(warps), 0,
                          _p(u1), _p(u2))
    elif method == 8:
        chi_a = f32(chi).copy()
        L.ref_tvl2occ(_p(I0), _p(I1c), _p(Im1), _p(u), _p(chi_a), params_file.encode(), w, h, int(warps),
                      int(glb_iters), 0)
    else:
        raise ValueError(method)
    return u, chi_a


# ------------------------------------------------------------------ file I/O
def read_flo(path):
    with open(path, "rb") as f:
        magic = np.frombuffer(f.read(4), np.float32)[0]
        assert magic == 202021.25, "bad .flo magic"
        w, h = np.frombuffer(f.read(8), np.int32)
        d = np.frombuffer(f.read(int(w) * int(h) * 8), np.float32).reshape(h, w, 2)
    return np.ascontiguousarray(d.transpose(2, 0, 1))


def write_flo(path, u):
    _, h, w = u.shape
    with open(path, "wb") as f:
        np.array([202021.25], np.float32).tofile(f)
        np.array([w, h], np.int32).tofile(f)
        np.ascontiguousarray(u.transpose(1, 2, 0), dtype=np.float32).tofile(f)


def write_ppm(path, rgb_u8):
    """rgb_u8: (h,w,3) uint8"""
    h, w, _ = rgb_u8.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgb_u8, np.uint8).tobytes())


def read_image_planar(path):
    """-> (pd,h,w) float32 0..255, like iio_read_image_float_split"""
    from PIL import Image
    im = np.asarray(Image.open(path))
    if im.ndim == 2:
        im = im[:, :, None]
    return np.ascontiguousarray(im[:, :, :3].transpose(2, 0, 1), dtype=np.float32)
