"""faldoi-ipol_b200: B200-native global_faldoi (FALDOI's global variational minimisation).

Thin ctypes layer over the C ABI in include/faldoi_gpu.h (libfaldoi_gpu.so, built
in-tree by build.py).  The function names mirror the reference's solver entry
points (src/global_faldoi.cpp: tvl2OF :556, nltvl1_PD :1177, tvcsad_PD :1449,
nltvcsad_PD :1642; src/tvl2_model_occ.cpp: guided_tvl2coupled_occ :492) so tests
read like calls into the reference.  There is no CPU fallback: if the library or
a B200 is missing, calls raise.

The directory name contains a hyphen; import it with
    importlib.import_module("faldoi-ipol_b200")
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FALDOI_GPU_LIB", os.path.join(HERE, "libfaldoi_gpu.so"))  # override only for kernel-variant experiments
MAX_WARPS = 64

M_TVL1, M_TVL1_W, M_NLTVL1, M_NLTVL1_W, M_TVCSAD, M_TVCSAD_W, M_NLTVCSAD, M_NLTVCSAD_W, M_TVL1_OCC = range(9)
METHOD_NAMES = {0: "tvl2", 1: "tvl2_w", 2: "nltv", 3: "nltv_w", 4: "tvcsad", 5: "tvcsad_w", 6: "nltvcsad",
                7: "nltvcsad_w", 8: "tvl2_occ"}


class FaldoiError(RuntimeError):
    pass


class Params(C.Structure):
    """faldoi_params (include/faldoi_gpu.h) == the scalars of `Parameters` (src/energy_structures.h:60-86)."""
    _fields_ = [("method", C.c_int), ("warps", C.c_int), ("max_iters", C.c_int)] + [
        (k, C.c_float) for k in ("lambda_", "theta", "tau", "beta", "alpha", "tau_u", "tau_eta", "tau_chi", "mu", "tol")]


class Log(C.Structure):
    _fields_ = [("iters", C.c_int * MAX_WARPS), ("err", C.c_float * MAX_WARPS)]


_fp = C.POINTER(C.c_float)
_lib = None

EXPORTS = [
    "faldoi_default_params", "faldoi_params_from_file", "faldoi_last_error", "faldoi_device_count",
    "faldoi_solver_create", "faldoi_solver_destroy", "faldoi_solver_upload", "faldoi_solver_run",
    "faldoi_solver_sync", "faldoi_solver_download", "faldoi_solver_last_run_ms", "faldoi_solver_last_iter_ms",
    "faldoi_solver_last_launches",
    "faldoi_solver_stream", "faldoi_solver_device_flow", "faldoi_global_solve", "faldoi_tvl2OF", "faldoi_tvcsad_PD",
    "faldoi_nltvl1_PD", "faldoi_nltvcsad_PD", "faldoi_guided_tvl2coupled_occ", "faldoi_centered_gradient",
    "faldoi_bicubic_warp", "faldoi_stripe_rows", "faldoi_stripes_create", "faldoi_stripes_destroy",
    "faldoi_stripes_upload", "faldoi_stripes_run", "faldoi_stripes_download", "faldoi_stripes_last_run_ms",
    "faldoi_stripes_last_launches", "faldoi_selftest_division", "faldoi_solver_upload_raw",
    "faldoi_solver_download_frames", "faldoi_global_solve_raw", "faldoi_solver_set_nltv_fast",
    "faldoi_solver_upload_xi", "faldoi_solver_download_xi", "faldoi_pinned_alloc", "faldoi_pinned_free",
]


def lib():
    """Load libfaldoi_gpu.so; never falls back to anything else."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FaldoiError("%s not built: run `python faldoi-ipol_b200/build.py` (needs nvcc)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.faldoi_last_error.restype = C.c_char_p
        L.faldoi_solver_last_run_ms.restype = C.c_float
        L.faldoi_solver_last_launches.restype = C.c_longlong
        L.faldoi_solver_last_iter_ms.restype = C.c_float
        L.faldoi_solver_last_iter_ms.argtypes = [C.c_void_p]
        L.faldoi_solver_stream.restype = C.c_void_p
        L.faldoi_solver_device_flow.restype = C.c_void_p
        L.faldoi_solver_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.faldoi_solver_destroy.argtypes = [C.c_void_p]
        L.faldoi_solver_upload.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.faldoi_solver_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int]
        L.faldoi_solver_sync.argtypes = [C.c_void_p]
        L.faldoi_solver_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Log)]
        L.faldoi_solver_last_run_ms.argtypes = [C.c_void_p]
        L.faldoi_solver_last_launches.argtypes = [C.c_void_p]
        L.faldoi_solver_stream.argtypes = [C.c_void_p]
        L.faldoi_solver_device_flow.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.faldoi_global_solve.argtypes = [C.c_int, C.POINTER(Params), C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.POINTER(Log)]
        vp, f, i = C.c_void_p, C.c_float, C.c_int
        L.faldoi_default_params.argtypes = [i, i, C.POINTER(Params)]
        L.faldoi_params_from_file.argtypes = [C.c_char_p, i, i, C.POINTER(Params)]
        L.faldoi_tvl2OF.argtypes = [vp] * 8 + [f] * 4 + [i] * 4
        L.faldoi_tvcsad_PD.argtypes = [vp] * 6 + [f] * 4 + [i] * 4 + [vp] * 2
        L.faldoi_nltvl1_PD.argtypes = [vp] * 3 + [i] + [f] * 3 + [i] * 4 + [vp] * 2
        L.faldoi_nltvcsad_PD.argtypes = [vp] * 3 + [i] + [f] * 3 + [i] * 4 + [vp] * 2
        L.faldoi_guided_tvl2coupled_occ.argtypes = [vp] * 6 + [C.POINTER(Params), i, i, i]
        L.faldoi_centered_gradient.argtypes = [i] + [vp] * 3 + [i, i]
        L.faldoi_bicubic_warp.argtypes = [i] + [vp] * 4 + [i, i, i]
        L.faldoi_stripe_rows.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
        L.faldoi_stripes_create.argtypes = [C.POINTER(vp), i, C.POINTER(i), i, i, i]
        L.faldoi_stripes_destroy.argtypes = [vp]
        L.faldoi_stripes_upload.argtypes = [vp, vp, vp, vp]
        L.faldoi_stripes_run.argtypes = [vp, C.POINTER(Params)]
        L.faldoi_stripes_download.argtypes = [vp, vp, C.POINTER(Log)]
        L.faldoi_stripes_last_run_ms.argtypes = [vp]
        L.faldoi_stripes_last_run_ms.restype = C.c_float
        L.faldoi_stripes_last_launches.argtypes = [vp]
        L.faldoi_stripes_last_launches.restype = C.c_longlong
        L.faldoi_selftest_division.argtypes = [i, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
        L.faldoi_solver_upload_raw.argtypes = [vp, i, vp, vp, vp, i, vp, vp]
        L.faldoi_solver_download_frames.argtypes = [vp, i, vp, vp, vp, vp]
        L.faldoi_global_solve_raw.argtypes = [i, C.POINTER(Params), i, i, i, vp, vp, vp, vp, vp, C.POINTER(Log)]
        L.faldoi_solver_set_nltv_fast.argtypes = [vp, i]
        L.faldoi_solver_upload_xi.argtypes = [vp, i, vp, vp, vp, vp]
        L.faldoi_solver_download_xi.argtypes = [vp, i, vp, vp, vp, vp]
        L.faldoi_pinned_alloc.argtypes = [C.c_size_t]
        L.faldoi_pinned_alloc.restype = vp
        L.faldoi_pinned_free.argtypes = [vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise FaldoiError("libfaldoi_gpu error %d: %s" % (rc, lib().faldoi_last_error().decode()))


def _ptr(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "planar float32 C-contiguous arrays only"
    return a.ctypes.data


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def default_params(method=M_TVL1, glb_iters=400, warps=5, params_file=None):
    """init_params + main()'s per-method overrides (src/global_faldoi.cpp:2022, 2138-2156)."""
    p = Params()
    if params_file:
        _check(lib().faldoi_params_from_file(params_file.encode(), int(method), int(glb_iters), C.byref(p)))
    else:
        _check(lib().faldoi_default_params(int(method), int(glb_iters), C.byref(p)))
    p.warps = int(warps)
    return p


def device_count():
    return lib().faldoi_device_count()


class Solver:
    """Batched handle: `batch` pairs of w x h resident in HBM on one device (faldoi_solver_*)."""

    def __init__(self, w, h, method=M_TVL1, batch=1, device=0):
        self._h = C.c_void_p()
        self.w, self.h, self.method, self.batch, self.device = w, h, method, batch, device
        _check(lib().faldoi_solver_create(C.byref(self._h), device, w, h, method, batch))

    def close(self):
        if self._h:
            lib().faldoi_solver_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, slot, I0, I1, u, Im1=None, lab=None, chi=None):
        arrs = [_f32(x) for x in (I0, I1, Im1, lab, u, chi)]
        _check(lib().faldoi_solver_upload(self._h, slot, *[_ptr(x) for x in arrs]))
        _check(lib().faldoi_solver_sync(self._h))  # host arrays may be temporaries

    def upload_raw(self, slot, i0, i1, im1, u, chi=None):
        """Raw planar frames (pd,h,w) 0..255; preprocessing runs on the device (faldoi_solver_upload_raw)."""
        arrs = [_f32(x) for x in (i0, i1, im1, u, chi)]
        _check(lib().faldoi_solver_upload_raw(self._h, slot, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), arrs[0].shape[0],
                                              _ptr(arrs[3]), _ptr(arrs[4])))
        _check(lib().faldoi_solver_sync(self._h))

    def download_frames(self, slot):
        """(I0n, I1n, Im1n or None, lab or None) as preprocessed on the device."""
        I0n, I1n = np.empty((self.h, self.w), np.float32), np.empty((self.h, self.w), np.float32)
        Im1n = np.empty((self.h, self.w), np.float32) if self.method == M_TVL1_OCC else None
        lab = np.empty((3, self.h, self.w), np.float32) if self.method in (2, 3, 6, 7) else None
        _check(lib().faldoi_solver_download_frames(self._h, slot, _ptr(I0n), _ptr(I1n), _ptr(Im1n), _ptr(lab)))
        return I0n, I1n, Im1n, lab

    def upload_ptrs(self, slot, I0, I1, u, Im1=0, lab=0, chi=0):
        """Raw host pointers (e.g. pinned torch tensors' data_ptr()); asynchronous."""
        _check(lib().faldoi_solver_upload(self._h, slot, I0, I1, Im1 or None, lab or None, u, chi or None))

    def set_nltv_fast(self, fast=True):
        """NLTV handles: opt into the approximate arithmetic mode (faldoi_solver_set_nltv_fast); not bit-exact."""
        _check(lib().faldoi_solver_set_nltv_fast(self._h, int(bool(fast))))

    def run(self, params, npairs=None):
        _check(lib().faldoi_solver_run(self._h, C.byref(params), self.batch if npairs is None else npairs))

    def sync(self):
        _check(lib().faldoi_solver_sync(self._h))

    def download(self, slot):
        u = np.empty((2, self.h, self.w), np.float32)
        chi = np.empty((self.h, self.w), np.float32) if self.method == M_TVL1_OCC else None
        log = Log()
        _check(lib().faldoi_solver_download(self._h, slot, _ptr(u), _ptr(chi), C.byref(log)))
        return u, chi, log

    def download_ptr(self, slot, u_ptr, chi_ptr=0):
        _check(lib().faldoi_solver_download(self._h, slot, u_ptr, chi_ptr or None, None))

    @property
    def last_run_ms(self):
        return lib().faldoi_solver_last_run_ms(self._h)

    @property
    def last_iter_ms(self):
        return lib().faldoi_solver_last_iter_ms(self._h)

    @property
    def last_launches(self):
        return lib().faldoi_solver_last_launches(self._h)

    @property
    def stream(self):
        return lib().faldoi_solver_stream(self._h)


def stripe_rows(h, nstripes, k):
    """Rows [row0, row1) of a frame of height h owned by stripe k of nstripes (faldoi_stripe_rows)."""
    r0, r1 = C.c_int(), C.c_int()
    _check(lib().faldoi_stripe_rows(h, nstripes, k, C.byref(r0), C.byref(r1)))
    return r0.value, r1.value


class Stripes:
    """One large frame pair cut into row stripes over several GPUs (faldoi_stripes_*), TVL2."""

    def __init__(self, w, h, devices, method=M_TVL1):
        self._h = C.c_void_p()
        self.w, self.h, self.devices = w, h, list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        _check(lib().faldoi_stripes_create(C.byref(self._h), len(self.devices), arr, w, h, method))

    def close(self):
        if self._h:
            lib().faldoi_stripes_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, I0, I1, u):
        arrs = [_f32(x) for x in (I0, I1, u)]
        _check(lib().faldoi_stripes_upload(self._h, *[_ptr(x) for x in arrs]))

    def upload_ptrs(self, I0, I1, u):
        _check(lib().faldoi_stripes_upload(self._h, I0, I1, u))

    def run(self, params):
        _check(lib().faldoi_stripes_run(self._h, C.byref(params)))

    def download(self):
        u = np.empty((2, self.h, self.w), np.float32)
        log = Log()
        _check(lib().faldoi_stripes_download(self._h, _ptr(u), C.byref(log)))
        return u, log

    def download_ptr(self, u_ptr):
        _check(lib().faldoi_stripes_download(self._h, u_ptr, None))

    @property
    def last_run_ms(self):
        return lib().faldoi_stripes_last_run_ms(self._h)

    @property
    def last_launches(self):
        return lib().faldoi_stripes_last_launches(self._h)


def global_solve(method, I0, I1, u, Im1=None, lab=None, chi=None, params=None, warps=5, glb_iters=400, device=0):
    """The call the host global_faldoi makes after preprocessing (faldoi_global_solve).
    Returns (u_out (2,h,w), chi_out or None, iters[warps], err[warps])."""
    h, w = I0.shape
    p = params or default_params(method, glb_iters, warps)
    u = _f32(u).copy()
    chi_a = _f32(chi).copy() if chi is not None else None
    log = Log()
    arrs = [_f32(I0), _f32(I1), _f32(Im1), _f32(lab)]
    _check(lib().faldoi_global_solve(device, C.byref(p), w, h, *[_ptr(x) for x in arrs], _ptr(u), _ptr(chi_a), C.byref(log)))
    return u, chi_a, list(log.iters[:p.warps]), list(log.err[:p.warps])


def global_solve_raw(method, i0, i1, im1, u, chi=None, params=None, warps=5, glb_iters=400, device=0):
    """global_faldoi from raw planar frames (pd,h,w) 0..255: preprocessing + solve on the device."""
    pd, h, w = i0.shape
    p = params or default_params(method, glb_iters, warps)
    u = _f32(u).copy()
    chi_a = _f32(chi).copy() if chi is not None else None
    log = Log()
    arrs = [_f32(i0), _f32(i1), _f32(im1)]
    _check(lib().faldoi_global_solve_raw(device, C.byref(p), w, h, pd, *[_ptr(x) for x in arrs], _ptr(u), _ptr(chi_a), C.byref(log)))
    return u, chi_a, list(log.iters[:p.warps]), list(log.err[:p.warps])


# ---- mirrors of the reference's solver signatures (in-place on u1,u2) --------
def tvl2OF(I0, I1, u1, u2, xi11, xi12, xi21, xi22, lambda_, theta, tau, tol_OF, nx, ny, warps, verbose):
    f = C.c_float
    _check(lib().faldoi_tvl2OF(_ptr(I0), _ptr(I1), _ptr(u1), _ptr(u2), _ptr(xi11), _ptr(xi12), _ptr(xi21), _ptr(xi22),
                               f(lambda_), f(theta), f(tau), f(tol_OF), int(nx), int(ny), int(warps), int(verbose)))


def tvcsad_PD(I0, I1, xi11, xi12, xi21, xi22, lambda_, theta, tau, tol_OF, nx, ny, warps, verbose, u1, u2):
    f = C.c_float
    _check(lib().faldoi_tvcsad_PD(_ptr(I0), _ptr(I1), _ptr(xi11), _ptr(xi12), _ptr(xi21), _ptr(xi22), f(lambda_),
                                  f(theta), f(tau), f(tol_OF), int(nx), int(ny), int(warps), int(verbose), _ptr(u1), _ptr(u2)))


def nltvl1_PD(I0, I1, a, pd, lambda_, theta, tau, w, h, warps, verbose, u1, u2):
    f = C.c_float
    _check(lib().faldoi_nltvl1_PD(_ptr(I0), _ptr(I1), _ptr(a), int(pd), f(lambda_), f(theta), f(tau), int(w), int(h),
                                  int(warps), int(verbose), _ptr(u1), _ptr(u2)))


def nltvcsad_PD(I0, I1, a, pd, lambda_, theta, tau, w, h, warps, verbose, u1, u2):
    f = C.c_float
    _check(lib().faldoi_nltvcsad_PD(_ptr(I0), _ptr(I1), _ptr(a), int(pd), f(lambda_), f(theta), f(tau), int(w), int(h),
                                    int(warps), int(verbose), _ptr(u1), _ptr(u2)))


def guided_tvl2coupled_occ(I0, I1, I_1, u1, u2, chi, params, nx, ny, verbose=0):
    _check(lib().faldoi_guided_tvl2coupled_occ(_ptr(I0), _ptr(I1), _ptr(I_1), _ptr(u1), _ptr(u2), _ptr(chi),
                                               C.byref(params), int(nx), int(ny), int(verbose)))


def selftest_division(n=1 << 28, seed=1, device=0):
    """Mismatches between the shared-reciprocal division and IEEE division over n operand pairs (x4 quotients)."""
    bad = C.c_ulonglong(0)
    _check(lib().faldoi_selftest_division(device, n, seed, C.byref(bad)))
    return bad.value


def centered_gradient(f, device=0):
    h, w = f.shape
    dx, dy = np.empty_like(f), np.empty_like(f)
    _check(lib().faldoi_centered_gradient(device, _ptr(f), _ptr(dx), _ptr(dy), w, h))
    return dx, dy


def bicubic_interpolation_warp(img, u, v, border_out, device=0):
    h, w = img.shape
    out = np.empty_like(img)
    _check(lib().faldoi_bicubic_warp(device, _ptr(img), _ptr(u), _ptr(v), _ptr(out), w, h, int(bool(border_out))))
    return out
