"""CPU: oracle primitives against the compiled reference (oracle/_ref), on random
inputs including the boundary cases.  Skipped where the reference build is absent."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L(po):
    if not po.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference): make -C oracle ref")
    return po.ref()


@pytest.mark.parametrize("w,h", [(2, 2), (5, 3), (17, 9), (64, 33)])
def test_stencils(po, L, w, h):
    rng = np.random.default_rng(w * 100 + h)
    f, a, b = [rng.standard_normal((h, w)).astype(np.float32) for _ in range(3)]
    for ours, theirs, args in ((po.o_centered_gradient, L.ref_centered_gradient, (f,)),
                               (po.o_forward_gradient, L.ref_forward_gradient, (f,))):
        rx, ry = np.empty_like(f), np.empty_like(f)
        theirs(po._p(f), po._p(rx), po._p(ry), w, h)
        ox, oy = ours(*args)
        assert np.array_equal(ox, rx) and np.array_equal(oy, ry)
    rd = np.empty_like(f)
    L.ref_divergence(po._p(a), po._p(b), po._p(rd), w, h)
    assert np.array_equal(po.o_divergence(a, b), rd)


@pytest.mark.parametrize("border_out", [0, 1])
def test_bicubic_warp(po, L, border_out):
    rng = np.random.default_rng(7)
    w, h = 37, 29
    img = rng.random((h, w)).astype(np.float32)
    # flows that leave the image on every side and cross zero (sign(uu) quirk, negative truncation)
    u = (rng.standard_normal((h, w)) * 6).astype(np.float32)
    v = (rng.standard_normal((h, w)) * 6).astype(np.float32)
    u[0, :5] = -0.5
    v[:5, 0] = -0.25
    r = np.empty_like(img)
    L.ref_bicubic_warp(po._p(img), po._p(u), po._p(v), po._p(r), w, h, border_out)
    o = po.o_bicubic_warp(img, u, v, border_out)
    assert np.array_equal(o, r)


def test_params_defaults(po, L):
    out = (C.c_float * 10)()
    L.ref_init_params(b"", out)
    p = po.default_params()
    ours = [p.lambda_, p.theta, p.tau, p.beta, p.alpha, p.tau_u, p.tau_eta, p.tau_chi, p.mu, p.tol]
    assert list(out) == ours
