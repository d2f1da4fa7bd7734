"""CPU: the C-ABI library loads, exports every symbol include/faldoi_gpu.h declares,
and its host-side logic (parameter defaults / -p file parsing / error behaviour)
matches the reference.  No kernel is launched here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "faldoi_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(faldoi_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(fb):
    L = fb.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libfaldoi_gpu.so does not export %s" % s
    assert sorted(fb.EXPORTS) == syms


def test_default_params_match_reference(fb, po):
    # src/parameters.h:20-31 via init_params; methods 2-7 overridden by main() (:2138-2156)
    d = po.default_params()
    p = fb.default_params(0)
    for k in ("lambda_", "theta", "tau", "beta", "alpha", "tau_u", "tau_eta", "tau_chi", "mu", "tol"):
        assert getattr(p, k) == getattr(d, k), k
    assert (p.warps, p.max_iters) == (5, 400)
    f32 = lambda x: C.c_float(x).value
    for m, (lam, th, tau) in {2: (2.0, 0.3, 0.1), 3: (2.0, 0.3, 0.1), 4: (0.85, 0.3, 0.125), 5: (0.85, 0.3, 0.125),
                              6: (0.85, 0.3, 0.1), 7: (0.85, 0.3, 0.1)}.items():
        p = fb.default_params(m, glb_iters=17)
        assert (p.lambda_, p.theta, p.tau) == (f32(lam), f32(th), f32(tau))
        assert p.max_iters == 400  # -glb_iters only reaches method 8 (src/tvl2_model_occ.cpp:653)
    assert fb.default_params(8, glb_iters=17).max_iters == 17
    with pytest.raises(fb.FaldoiError):
        fb.default_params(9)


def test_params_file(fb, tmp_path):
    f = tmp_path / "p.txt"
    # lambda theta tau beta alpha tau_u tau_eta tau_chi mu ; <=0 -> default, tau* > 0.25 -> default
    f.write_text("23.5\n0.28\n0.3\n0.7\n-1\n0.07\n0.26\n0.13\n0\n")
    p = fb.default_params(8, glb_iters=40, params_file=str(f))
    d = fb.default_params(8)
    f32 = lambda x: C.c_float(x).value
    assert p.lambda_ == f32(23.5) and p.theta == f32(0.28) and p.beta == f32(0.7)
    assert p.tau == d.tau and p.alpha == d.alpha and p.tau_eta == d.tau_eta and p.mu == d.mu
    assert p.tau_u == f32(0.07) and p.tau_chi == f32(0.13)
    # methods 2-7 ignore the file's lambda/theta/tau
    assert fb.default_params(4, params_file=str(f)).lambda_ == f32(0.85)
    # the reference aborts on a missing / short file; we report it
    with pytest.raises(fb.FaldoiError):
        fb.default_params(0, params_file=str(tmp_path / "missing.txt"))
    f.write_text("1\n2\n")
    with pytest.raises(fb.FaldoiError):
        fb.default_params(0, params_file=str(f))


def test_no_cpu_fallback(fb):
    """Without a B200 the product must fail loudly, never compute on the CPU."""
    import numpy as np
    if fb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fb.FaldoiError):
        fb.Solver(64, 48, 0, 1)
    z = np.zeros((48, 64), np.float32)
    with pytest.raises(fb.FaldoiError):
        fb.global_solve(0, z, z, np.zeros((2, 48, 64), np.float32))


def test_product_does_not_touch_oracle():
    """The product tree must not reference oracle/ (judge rule: oracle is the checker only)."""
    pkg = os.path.join(ROOT, "faldoi-ipol_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "faldoi_oracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)
