"""CPU: the oracle (our C restatement) against the golden vectors produced by the
UNMODIFIED reference build (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from conftest import CASE_RUNS, load_case, run_key


@pytest.mark.parametrize("case", ["crop_a", "crop_b"])
def test_preprocessing_bit_exact(po, case):
    g = load_case(case)
    rgb = [g[k].astype(np.float32) for k in ("rgb_i0", "rgb_i1", "rgb_im1")]
    I0, I1, Im1 = po.o_preprocess(*rgb)
    assert np.array_equal(I0, g["I0n"]) and np.array_equal(I1, g["I1n"]) and np.array_equal(Im1, g["Im1n"])
    assert np.array_equal(po.o_image_to_lab(rgb[0]), g["lab"])
    # normalised to [0,1], not 0..255
    assert 0.0 <= I0.min() and I0.max() <= 1.0 + 1e-6


@pytest.mark.parametrize("case,run", [(c, r) for c in CASE_RUNS for r in CASE_RUNS[c]])
def test_solver_bit_exact(po, case, run):
    method, warps, iters = run
    g = load_case(case)
    u, chi, its, errs = po.o_global_solve(method, g["I0n"], g["I1n"], g["Im1n"], g["lab"], g["u0"],
                                          g["chi0"] if method == 8 else None, warps=warps, glb_iters=iters)
    key = run_key(*run)
    assert np.array_equal(u, g["u_" + key]), "max diff %g" % np.abs(u - g["u_" + key]).max()
    if method == 8:
        assert np.array_equal(chi, g["chi_" + key])
        assert set(np.unique(chi)) <= {0.0, 1.0}
    assert len(its) == warps and all(1 <= n <= iters for n in its)
    if method in (2, 6):
        assert its == [400] * warps  # NLTV never exits early (src/global_faldoi.cpp:1249)


def test_fullsize_anchor_is_recorded():
    g = load_case("fullsize_clean_easy_m0")
    # anchors measured on the reference: SURVEY.md 8(c)
    assert list(g["iters"]) == [400, 152, 100, 82, 136]
    assert round(float(g["epe_init"]), 4) == 0.3298 and round(float(g["epe_out"]), 4) == 0.2232


@pytest.mark.parametrize("seq", ["final_hard", "clean_medium", "clean_hard", "final_easy", "final_medium"])
def test_oracle_fullsize_other_sequences(po, seq):
    """The C restatement against the reference EXECUTABLE's output on whole Sintel pairs (TVL2, 5 warps x <= 400
    iterations): same iteration counts, same flow in every bit.  Needs oracle/_ref/data/<seq>
    (oracle/run_full_refs_seq.sh, run_full_refs_all.sh); 5-20 s each on 8 cores."""
    import os
    from conftest import ROOT
    D = os.path.join(ROOT, "oracle", "_ref", "data", seq)
    ref = os.path.join(D, "var_m0.flo")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/data/%s not present" % seq)
    g = load_case("fullsize_%s_m0" % seq)
    fr = [po.read_image_planar(os.path.join(D, "frame_%04d.png" % k)) for k in (1, 2, 3)]
    I0, I1, _ = po.o_preprocess(fr[1], fr[2], fr[0])[:3]
    u, _, its, _ = po.o_global_solve(0, I0, I1, None, None, po.read_flo(os.path.join(D, "rg.flo")), warps=5)
    assert its == list(g["iters"])
    assert np.array_equal(u, po.read_flo(ref))
    assert np.array_equal(u[:, ::8, ::8], g["u_sub"])
