import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def po():
    """oracle bindings (test infrastructure)"""
    import pyoracle
    pyoracle.oracle()
    return pyoracle


@pytest.fixture(scope="session")
def fb():
    """the product package (ctypes over libfaldoi_gpu.so)"""
    return importlib.import_module("faldoi-ipol_b200")


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


# (method, warps, glb_iters) runs stored per golden case by tests/golden/make_golden.py
CASE_RUNS = {
    "crop_a": [(0, 5, 400), (4, 2, 400), (2, 2, 400), (6, 2, 400), (8, 2, 30)],
    "crop_b": [(0, 3, 400), (4, 1, 400), (2, 1, 400), (6, 1, 400), (8, 1, 12)],
}


def run_key(method, warps, iters):
    return "m%d_w%d" % (method, warps) + ("_i%d" % iters if method == 8 else "")


def synthetic_pair(w, h, seed, max_flow=4.0):
    """Seeded synthetic pair in the spirit of SURVEY.md 8(d): band-limited texture,
    smooth true flow, I1 / I-1 = I0 shifted along +/- flow (nearest+blur is enough for a
    solver test), init flow = truth + smooth noise.  Returns preprocessed-like gray
    frames in [0,1] plus a 3-channel 0..255 'rgb' of I0 for the Lab path."""
    rng = np.random.default_rng(seed)

    def smooth_noise(sigma):
        a = rng.standard_normal((h, w)).astype(np.float32)
        k = int(3 * sigma) | 1
        ker = np.exp(-0.5 * (np.arange(-k, k + 1) / sigma) ** 2).astype(np.float32)
        ker /= ker.sum()
        a = np.apply_along_axis(lambda r: np.convolve(np.pad(r, k, mode="reflect"), ker, "valid"), 1, a)
        a = np.apply_along_axis(lambda r: np.convolve(np.pad(r, k, mode="reflect"), ker, "valid"), 0, a)
        return a.astype(np.float32)

    tex = sum(smooth_noise(2.0 ** k) * (2.0 ** k) for k in range(4))
    tex = (tex - tex.min()) / (tex.max() - tex.min())
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    f1 = max_flow * 0.5 * (np.sin(2 * np.pi * xx / w + 0.3) + 0.5 * np.cos(2 * np.pi * yy / h))
    f2 = max_flow * 0.5 * (np.cos(2 * np.pi * xx / w * 1.5) - 0.5 * np.sin(2 * np.pi * yy / h + 0.7))

    def shift(img, s):
        xs = np.clip(np.rint(xx + s * f1), 0, w - 1).astype(np.int64)
        ys = np.clip(np.rint(yy + s * f2), 0, h - 1).astype(np.int64)
        return img[ys, xs]

    I0 = tex.astype(np.float32)
    I1 = shift(tex, -1.0).astype(np.float32)
    Im1 = shift(tex, 1.0).astype(np.float32)
    u0 = np.stack([f1 + 0.5 * smooth_noise(3.0) * 3, f2 + 0.5 * smooth_noise(3.0) * 3]).astype(np.float32)
    rgb = np.stack([I0 * 255, np.roll(I0, 3, 1) * 200 + 20, np.roll(I0, 5, 0) * 180 + 40]).astype(np.float32)
    return (np.ascontiguousarray(I0), np.ascontiguousarray(I1), np.ascontiguousarray(Im1), np.ascontiguousarray(u0),
            np.ascontiguousarray(rgb))
