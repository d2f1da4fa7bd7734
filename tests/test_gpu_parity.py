"""GPU: the sm_100a path (through the C ABI) against the oracle and the reference's
golden vectors.  Bars: TVL2, TV-CSAD and TVL2-OCC are bit-exact (fp32, same
operation order, no FMA); the NLTV models compute their 24 weights with the GPU's
exp and are held to the north star's tolerance (mean |du| <= 1e-3, max <= 1e-2 px)."""
import numpy as np
import pytest

from conftest import CASE_RUNS, load_case, run_key, synthetic_pair

pytestmark = pytest.mark.gpu

MEAN_TOL, MAX_TOL = 1e-3, 1e-2  # px, BASELINE.json north_star


def assert_flow(u, ref, exact):
    d = np.abs(u - ref)
    if exact:
        assert np.array_equal(u, ref), "not bit-exact: max |du| = %g, %d px differ" % (d.max(), (d > 0).sum())
    else:
        assert d.mean() <= MEAN_TOL and d.max() <= MAX_TOL, "mean %g max %g" % (d.mean(), d.max())


@pytest.mark.parametrize("case,run", [(c, r) for c in CASE_RUNS for r in CASE_RUNS[c]])
def test_golden(fb, po, case, run):
    method, warps, iters = run
    g = load_case(case)
    chi0 = g["chi0"] if method == 8 else None
    u, chi, its, errs = fb.global_solve(method, g["I0n"], g["I1n"], g["u0"], Im1=g["Im1n"], lab=g["lab"], chi=chi0,
                                        warps=warps, glb_iters=iters)
    key = run_key(*run)
    assert_flow(u, g["u_" + key], exact=True)  # every energy model is bit-identical to the reference
    _, ochi, oits, oerrs = po.o_global_solve(method, g["I0n"], g["I1n"], g["Im1n"], g["lab"], g["u0"], chi0,
                                             warps=warps, glb_iters=iters)
    assert its == oits
    if method in (0, 8):
        assert errs == oerrs  # max-norm error is order independent
    else:
        assert np.allclose(errs, oerrs, rtol=1e-4)
    if method == 8:
        assert np.array_equal(chi, g["chi_" + key])


@pytest.mark.parametrize("w,h", [(33, 20), (128, 32), (130, 67), (257, 9)])
@pytest.mark.parametrize("method", [0, 4])
def test_ragged_sizes_vs_oracle(fb, po, w, h, method):
    I0, I1, Im1, u0, _ = synthetic_pair(w, h, seed=w + h)
    u, _, its, _ = fb.global_solve(method, I0, I1, u0, warps=2)
    ou, _, oits, _ = po.o_global_solve(method, I0, I1, None, None, u0, warps=2)
    assert its == oits
    assert_flow(u, ou, exact=True)


@pytest.mark.parametrize("w,h", [(300, 131), (257, 9), (130, 67), (129, 5)])
@pytest.mark.parametrize("method", [2, 6, 8])
def test_ragged_sizes_vs_oracle_other_models(fb, po, w, h, method):
    """The NLTV / NLTV-CSAD / TVL2-OCC tile kernels on frames wider than one and two tiles (NLTV tiles are 128
    wide, OCC tiles 120) and not a multiple of anything: bit-exact against the oracle, chi included."""
    I0, I1, Im1, u0, rgb = synthetic_pair(w, h, seed=3 * w + h + method)
    lab = po.o_image_to_lab(rgb) if method in (2, 6) else None
    chi0 = (np.random.default_rng(w + h).random((h, w)) > 0.85).astype(np.float32) if method == 8 else None
    warps, iters = (1, 400) if method != 8 else (2, 8)
    u, chi, its, _ = fb.global_solve(method, I0, I1, u0, Im1=Im1 if method == 8 else None, lab=lab, chi=chi0, warps=warps,
                                     glb_iters=iters)
    ou, ochi, oits, _ = po.o_global_solve(method, I0, I1, Im1, lab, u0, chi0, warps=warps, glb_iters=iters)
    assert its == oits
    assert_flow(u, ou, exact=True)
    if method == 8:
        assert np.array_equal(chi, ochi)


def test_4k_tvl2_vs_oracle(fb, po):
    """BASELINE configs[4]'s large shape: one 3840x2160 TVL2 warp against the oracle on the host cores (the frame
    is 32 x 240 tiles of the two-iteration kernel; 41 planes of 33 MB are far beyond L2)."""
    I0, I1, _, u0, _ = synthetic_pair(3840, 2160, seed=4)
    u, _, its, errs = fb.global_solve(0, I0, I1, u0, warps=1)
    ou, _, oits, oerrs = po.o_global_solve(0, I0, I1, None, None, u0, warps=1)
    assert its == oits and errs == oerrs
    assert np.array_equal(u, ou)


def test_batch_equals_single(fb, po):
    """Pairs of one batch are independent problems with their own exit iteration."""
    w, h, B = 96, 64, 5
    pairs = [synthetic_pair(w, h, seed=100 + k, max_flow=1.0 + k) for k in range(B)]
    s = fb.Solver(w, h, 0, B)
    for k, (I0, I1, _, u0, _) in enumerate(pairs):
        s.upload(k, I0, I1, u0)
    p = fb.default_params(0, warps=3)
    s.run(p)
    iters_seen = set()
    for k, (I0, I1, _, u0, _) in enumerate(pairs):
        u, _, log = s.download(k)
        ou, _, oits, _ = po.o_global_solve(0, I0, I1, None, None, u0, warps=3)
        assert list(log.iters[:3]) == oits
        assert np.array_equal(u, ou)
        iters_seen.add(tuple(oits))
    assert len(iters_seen) > 1, "test needs pairs with different exit iterations"
    assert s.last_launches > 3 * 50  # iteration launches really ran (early termination stops enqueuing once all pairs exit)
    s.close()


def test_reference_signature_mirrors(fb, po):
    g = load_case("crop_b")
    h, w = g["I0n"].shape
    u1, u2 = g["u0"][0].copy(), g["u0"][1].copy()
    xi = [np.zeros((h, w), np.float32) for _ in range(4)]
    fb.tvl2OF(g["I0n"], g["I1n"].copy(), u1, u2, *xi, 40.0, 0.3, 0.125, 0.01, w, h, 3, 0)
    assert np.array_equal(np.stack([u1, u2]), g["u_m0_w3"])
    u1, u2 = g["u0"][0].copy(), g["u0"][1].copy()
    xi = [np.zeros((h, w), np.float32) for _ in range(4)]  # in/out arguments: the call above left its final duals in them
    fb.tvcsad_PD(g["I0n"], g["I1n"].copy(), *xi, 0.85, 0.3, 0.125, 0.01, w, h, 1, 0, u1, u2)
    assert np.array_equal(np.stack([u1, u2]), g["u_m4_w1"])
    u1, u2 = g["u0"][0].copy(), g["u0"][1].copy()
    fb.nltvl1_PD(g["I0n"], g["I1n"].copy(), g["lab"].copy(), 3, 2.0, 0.3, 0.1, w, h, 1, 0, u1, u2)
    assert_flow(np.stack([u1, u2]), g["u_m2_w1"], exact=False)
    u1, u2, chi = g["u0"][0].copy(), g["u0"][1].copy(), g["chi0"].copy()
    fb.guided_tvl2coupled_occ(g["I0n"], g["I1n"], g["Im1n"], u1, u2, chi, fb.default_params(8, 12, 1), w, h)
    assert np.array_equal(np.stack([u1, u2]), g["u_m8_w1_i12"]) and np.array_equal(chi, g["chi_m8_w1_i12"])


def test_mirrors_read_and_return_the_duals(fb, po):
    """tvl2OF's xi arguments are in/out (src/global_faldoi.cpp:556-573): two 1-warp calls that hand the duals over
    must equal one 2-warp call (the reference carries xi across warps), and the returned duals equal the oracle's."""
    g = load_case("crop_b")
    h, w = g["I0n"].shape
    u1, u2 = g["u0"][0].copy(), g["u0"][1].copy()
    xi = [np.zeros((h, w), np.float32) for _ in range(4)]
    for _ in range(2):
        fb.tvl2OF(g["I0n"], g["I1n"].copy(), u1, u2, *xi, 40.0, 0.3, 0.125, 0.01, w, h, 1, 0)
    ou, oxi, _, _ = po.o_tvl2(g["I0n"], g["I1n"], g["u0"], warps=2)
    assert np.array_equal(np.stack([u1, u2]), ou)
    assert np.array_equal(np.stack(xi), oxi)
    assert any(np.abs(x).max() > 0 for x in xi)


def test_nltv_second_run_continues_from_the_first(fb, po):
    """A second faldoi_solver_run on an NLTV handle without a fresh upload continues from the state the first one
    left (odd warps*max_iters leaves it in the other ping-pong set): 1 + 1 warps of 3 iterations == 2 warps."""
    g = load_case("crop_b")
    h, w = g["I0n"].shape
    p = fb.default_params(2, warps=1)
    p.max_iters = 3
    s = fb.Solver(w, h, 2, 1)
    s.upload(0, g["I0n"], g["I1n"], g["u0"], lab=g["lab"])
    s.run(p)
    s.run(p)
    u, _, _ = s.download(0)
    p2 = fb.default_params(2, warps=2)
    p2.max_iters = 3
    s.upload(0, g["I0n"], g["I1n"], g["u0"], lab=g["lab"])
    s.run(p2)
    u2, _, _ = s.download(0)
    s.close()
    assert np.array_equal(u, u2)


def test_primitives(fb, po):
    rng = np.random.default_rng(3)
    w, h = 75, 41
    img = rng.random((h, w)).astype(np.float32)
    dx, dy = fb.centered_gradient(img)
    ox, oy = po.o_centered_gradient(img)
    assert np.array_equal(dx, ox) and np.array_equal(dy, oy)
    u = (rng.standard_normal((h, w)) * 5).astype(np.float32)
    v = (rng.standard_normal((h, w)) * 5).astype(np.float32)
    for bo in (0, 1):
        assert np.array_equal(fb.bicubic_interpolation_warp(img, u, v, bo), po.o_bicubic_warp(img, u, v, bo))


def test_fullsize_tvl2_vs_oracle(fb, po):
    """BASELINE config 1 shape (1024x436), synthetic content; oracle on the host cores."""
    I0, I1, _, u0, _ = synthetic_pair(1024, 436, seed=1)
    u, _, its, errs = fb.global_solve(0, I0, I1, u0, warps=5)
    ou, _, oits, oerrs = po.o_global_solve(0, I0, I1, None, None, u0, warps=5)
    assert its == oits and errs == oerrs
    assert np.array_equal(u, ou)


def test_fullsize_reference_pair(fb, po):
    """The named real config: clean/easy Sintel pair with the local_faldoi flow as init.
    Inputs live in oracle/_ref/data (git-ignored, travels with the snapshot); anchors
    (iteration counts, EPE 0.2232) come from the committed golden file."""
    import os
    from conftest import ROOT
    D = os.path.join(ROOT, "oracle", "_ref", "data", "clean_easy")
    if not os.path.exists(os.path.join(D, "rg.flo")):
        pytest.skip("oracle/_ref/data/clean_easy not present (oracle/make_init_flow.sh)")
    fr = [po.read_image_planar(os.path.join(D, "frame_%04d.png" % k)) for k in (1, 2, 3)]
    I0, I1, Im1 = po.o_preprocess(fr[1], fr[2], fr[0])
    u0 = po.read_flo(os.path.join(D, "rg.flo"))
    u, _, its, _ = fb.global_solve(0, I0, I1, u0, warps=5)
    g = load_case("fullsize_clean_easy_m0")
    assert its == list(g["iters"])
    assert np.array_equal(u[:, ::8, ::8], g["u_sub"])
    gt = po.read_flo(os.path.join(D, "gt_frame_0002.flo"))
    epe = float(np.sqrt(((u - gt) ** 2).sum(0)).mean())
    assert round(epe, 3) == round(float(g["epe_out"]), 3) == 0.223
    if os.path.exists(os.path.join(D, "var_m0.flo")):
        assert np.array_equal(u, po.read_flo(os.path.join(D, "var_m0.flo")))


def test_shared_reciprocal_division_is_ieee(fb):
    """The 4-quotients-one-reciprocal division of the TVL2 dual projection must equal IEEE division
    bit for bit on its guarded range: 2^31 quotients, random and structured significands."""
    for seed in (1, 2):
        assert fb.selftest_division(1 << 28, seed) == 0


@pytest.mark.parametrize("case", ["crop_a", "crop_b"])
def test_device_preprocessing(fb, po, case):
    """upload_raw: gray / joint normalisation / Gaussian on the device are bit-identical to the reference's
    (golden I0n, I1n, Im1n); Lab goes through the device's pow/exp and is tolerance-level."""
    g = load_case(case)
    raw = [np.ascontiguousarray(g[k].astype(np.float32)) for k in ("rgb_i0", "rgb_i1", "rgb_im1")]
    _, h, w = raw[0].shape
    s = fb.Solver(w, h, 8, 1)
    s.upload_raw(0, raw[0], raw[1], raw[2], g["u0"], g["chi0"])
    I0n, I1n, Im1n, _ = s.download_frames(0)
    assert np.array_equal(I0n, g["I0n"]) and np.array_equal(I1n, g["I1n"]) and np.array_equal(Im1n, g["Im1n"])
    s.close()
    s = fb.Solver(w, h, 2, 1)
    s.upload_raw(0, raw[0], raw[1], raw[2], g["u0"])
    _, _, _, lab = s.download_frames(0)
    assert np.abs(lab - g["lab"]).max() <= 1e-4 * max(1.0, np.abs(g["lab"]).max())
    s.close()
    # gray single-channel input takes the pd == 1 path
    gray = [po.o_preprocess(*raw)[k] for k in range(3)]  # any (1,h,w) float planes will do as "raw gray" input
    s = fb.Solver(w, h, 0, 1)
    s.upload_raw(0, gray[0][None] * 255, gray[1][None] * 255, gray[2][None] * 255, g["u0"])
    a, b_, _, _ = s.download_frames(0)
    oa, ob, _ = po.o_preprocess(gray[0][None] * 255, gray[1][None] * 255, gray[2][None] * 255)
    assert np.array_equal(a, oa) and np.array_equal(b_, ob)
    s.close()


def test_global_solve_raw_matches_reference(fb):
    g = load_case("crop_b")
    raw = [np.ascontiguousarray(g[k].astype(np.float32)) for k in ("rgb_i0", "rgb_i1", "rgb_im1")]
    u, _, its, _ = fb.global_solve_raw(0, raw[0], raw[1], raw[2], g["u0"], warps=3)
    assert np.array_equal(u, g["u_m0_w3"])
    u, chi, _, _ = fb.global_solve_raw(8, raw[0], raw[1], raw[2], g["u0"], chi=g["chi0"], warps=1, glb_iters=12)
    assert np.array_equal(u, g["u_m8_w1_i12"]) and np.array_equal(chi, g["chi_m8_w1_i12"])
    u, _, _, _ = fb.global_solve_raw(6, raw[0], raw[1], raw[2], g["u0"], warps=1)
    assert_flow(u, g["u_m6_w1"], exact=False)


@pytest.mark.parametrize("max_iters,tol", [(1, 0.0), (2, 0.0), (3, 0.0), (7, 0.0), (400, 10.0), (400, 0.02)])
def test_exit_iteration_corner_cases(fb, po, max_iters, tol):
    """Two iterations run per launch: odd iteration caps, an exit after the very first iteration and
    exits that fall on the first / second half of a launch must all reproduce the reference loop."""
    I0, I1, _, u0, _ = synthetic_pair(150, 40, seed=11)
    p = fb.default_params(0, warps=3)
    p.max_iters, p.tol = max_iters, tol
    u, _, its, errs = fb.global_solve(0, I0, I1, u0, params=p)
    ou, _, oits, oerrs = po.o_tvl2(I0, I1, u0, tol=tol, warps=3, max_iter=max_iters)
    assert its == oits and errs == oerrs
    assert np.array_equal(u, ou)


@pytest.mark.parametrize("lam,theta,tau", [(23.4547299737, 0.283263949386, 0.125), (40.0, 0.25, 0.2), (15.0, 0.683018503834, 0.0739776273913)])
def test_custom_parameters_tvl2(fb, po, lam, theta, tau):
    """-p files change lambda/theta/tau for methods 0,1,8 (README 'best params'): the constant-divisor
    fast path is re-verified per theta on the device and must stay bit-exact."""
    I0, I1, _, u0, _ = synthetic_pair(200, 56, seed=21)
    p = fb.default_params(0, warps=2)
    p.lambda_, p.theta, p.tau = lam, theta, tau
    u, _, its, errs = fb.global_solve(0, I0, I1, u0, params=p)
    ou, _, oits, oerrs = po.o_tvl2(I0, I1, u0, lam=lam, theta=theta, tau=tau, warps=2)
    assert its == oits and errs == oerrs
    assert np.array_equal(u, ou)


def test_custom_parameters_occ(fb, po, tmp_path):
    """Method 8 reads all nine scalars of the -p file."""
    g = load_case("crop_b")
    f = tmp_path / "p.txt"
    f.write_text("23.4547299737\n0.283263949386\n0.125\n0.701945704713\n0.0706776435878\n0.0739776273913\n0.0839911992024\n0.134077646787\n1.4058686732\n")
    p = fb.default_params(8, glb_iters=6, warps=1, params_file=str(f))
    u, chi, its, _ = fb.global_solve(8, g["I0n"], g["I1n"], g["u0"], Im1=g["Im1n"], chi=g["chi0"], params=p)
    op = po.default_params()
    op.lambda_, op.theta, op.beta = p.lambda_, p.theta, p.beta
    ou, ochi, oits, _ = po.o_global_solve(8, g["I0n"], g["I1n"], g["Im1n"], None, g["u0"], g["chi0"], params=op, warps=1, glb_iters=6)
    assert its == oits
    assert np.array_equal(u, ou) and np.array_equal(chi, ochi)


def test_zero_warps_and_single_pixel_rows(fb, po):
    I0, I1, _, u0, _ = synthetic_pair(64, 5, seed=3)
    p = fb.default_params(0, warps=0)
    u, _, its, _ = fb.global_solve(0, I0, I1, u0, params=p)
    assert its == [] and np.array_equal(u, u0)
    u, _, its, _ = fb.global_solve(0, I0, I1, u0, warps=1)
    ou, _, oits, _ = po.o_global_solve(0, I0, I1, None, None, u0, warps=1)
    assert its == oits and np.array_equal(u, ou)


@pytest.mark.parametrize("method", [4, 7, 8])
def test_fullsize_reference_pair_other_models(fb, po, method):
    """BASELINE configs 2-4 at full size: the reference CLI's result on clean/easy (oracle/run_full_refs.sh,
    minutes of CPU each) vs one GPU solve.  TV-CSAD and TVL2-OCC bit for bit, NLTV-CSAD(-W) within the north
    star's tolerance; EPE against the Sintel ground truth equal to 3 decimals."""
    import os
    from conftest import ROOT, GOLDEN
    D = os.path.join(ROOT, "oracle", "_ref", "data", "clean_easy")
    gfile = os.path.join(GOLDEN, "fullsize_clean_easy_m%d.npz" % method)
    init = os.path.join(D, "rg_m8.flo" if method == 8 else "rg.flo")
    if not os.path.exists(gfile) or not os.path.exists(init):
        pytest.skip("full-size golden / inputs not present")
    g = dict(np.load(gfile))
    fr = [po.read_image_planar(os.path.join(D, "frame_%04d.png" % k)) for k in (1, 2, 3)]
    u0 = po.read_flo(init)
    chi0 = None
    if method == 8:
        from PIL import Image
        chi0 = np.asarray(Image.open(os.path.join(D, "rg_occ.png"))).astype(np.float32)
    u, chi, its, _ = fb.global_solve_raw(method, fr[1], fr[2], fr[0], u0, chi=chi0, warps=5, glb_iters=400)
    assert its == list(g["iters"])
    gt = po.read_flo(os.path.join(D, "gt_frame_0002.flo"))
    epe = float(np.sqrt(((u - gt) ** 2).sum(0)).mean())
    assert round(epe, 3) == round(float(g["epe_out"]), 3)
    if method == 7:
        # raw frames: gray / normalise / Gaussian / Lab all on the device (the Lab attenuation uses glibc's expf
        # algorithm; its cube roots use the device's double pow, which rounds to the same floats as glibc's on
        # this frame) -> the flow is bit-identical to the reference executable's, as with the reference's own
        # host-side preprocessing
        assert np.array_equal(u[:, ::8, ::8], g["u_sub"])
        full = os.path.join(D, "var_m7.flo")
        if os.path.exists(full):
            ref = po.read_flo(full)
            assert np.array_equal(u, ref), "max |du| = %g" % np.abs(u - ref).max()
            I0, I1, Im1 = po.o_preprocess(fr[1], fr[2], fr[0])[:3]
            uh, _, its_h, _ = fb.global_solve(method, I0, I1, u0, Im1=Im1, lab=po.o_image_to_lab(fr[1]), warps=5, glb_iters=400)
            assert its_h == list(g["iters"]) and np.array_equal(uh, ref)
    else:
        assert np.array_equal(u[:, ::8, ::8], g["u_sub"])
        full = os.path.join(D, "var_m%d.flo" % method)
        if os.path.exists(full):
            assert np.array_equal(u, po.read_flo(full))
    if method == 8:
        assert np.array_equal(chi[::4, ::4].astype(np.uint8), g["chi_sub"]) and int(chi.sum()) == int(g["chi_count"])


@pytest.mark.parametrize("method", [2, 6])
def test_nltv_fast_mode_within_tolerance(fb, po, method):
    """faldoi_solver_set_nltv_fast: approximate divisions and paired slot order, the opt-in throughput mode of the
    NLTV models -- held to the north star's tolerance on the goldens (the default mode is bit-exact, test_golden)."""
    for case in ("crop_a", "crop_b"):
        g = load_case(case)
        run = [r for r in CASE_RUNS[case] if r[0] == method][0]
        h, w = g["I0n"].shape
        s = fb.Solver(w, h, method, 1)
        s.set_nltv_fast(True)
        s.upload(0, g["I0n"], g["I1n"], g["u0"], lab=g["lab"])
        s.run(fb.default_params(method, run[2], run[1]))
        u, _, _ = s.download(0)
        s.close()
        ref = g["u_" + run_key(*run)]
        assert not np.array_equal(u, ref), "fast mode did not take effect"
        assert_flow(u, ref, exact=False)


@pytest.mark.parametrize("method,warps,iters", [(4, 1, 400), (2, 1, 400), (6, 1, 400), (8, 1, 6)])
def test_batch_equals_oracle_other_models(fb, po, method, warps, iters):
    """The batched handle with the other energy models: every slot of a batch of different pairs equals the
    oracle's solve of that pair bit for bit (tile staging, TMA maps and the CSAD / NLTV tables are all indexed
    by slot)."""
    w, h, B = 70, 37, 3  # ragged: not a multiple of the quad or of any tile
    pairs = [synthetic_pair(w, h, seed=300 + k, max_flow=1.0 + k) for k in range(B)]
    labs = [po.o_image_to_lab(p[4]) for p in pairs]
    chis = [(np.random.default_rng(k).random((h, w)) > 0.8).astype(np.float32) for k in range(B)]
    s = fb.Solver(w, h, method, B)
    for k, (I0, I1, Im1, u0, _) in enumerate(pairs):
        s.upload(k, I0, I1, u0, Im1=Im1 if method == 8 else None, lab=labs[k] if method in (2, 6) else None,
                 chi=chis[k] if method == 8 else None)
    p = fb.default_params(method, glb_iters=iters, warps=warps)
    s.run(p)
    for k, (I0, I1, Im1, u0, _) in enumerate(pairs):
        u, chi, log = s.download(k)
        ou, ochi, oits, _ = po.o_global_solve(method, I0, I1, Im1, labs[k], u0, chis[k] if method == 8 else None, warps=warps,
                                              glb_iters=iters)
        assert list(log.iters[:warps]) == oits
        assert np.array_equal(u, ou), "slot %d: max |du| = %g" % (k, np.abs(u - ou).max())
        if method == 8:
            assert np.array_equal(chi, ochi)
    s.close()


@pytest.mark.parametrize("tol", [0.05, 0.15, 0.3, 1.0])
def test_tvcsad_exit_test(fb, po, tol):
    """TV-CSAD leaves a warp when the mean squared update drops to tol^2 (src/global_faldoi.cpp:1543; the sum over
    ALL pixels, which is what the reference computes with one thread -- DESIGN.md on the racy err_D).  Larger
    tolerances make the exit happen early and at different iterations per warp."""
    I0, I1, _, u0, _ = synthetic_pair(140, 50, seed=21, max_flow=1.5)
    p = fb.default_params(4, warps=3)
    p.tol = tol
    u, _, its, errs = fb.global_solve(4, I0, I1, u0, params=p)
    op = po.default_params()
    op.tol = tol
    ou, _, oits, oerrs = po.o_global_solve(4, I0, I1, None, None, u0, params=op, warps=3)
    assert its == oits, (its, oits)
    assert np.array_equal(u, ou)
    assert np.allclose(errs, oerrs, rtol=1e-4)
    if tol >= 0.3:
        assert min(its) < 400, "the exit test was never met: %s" % its
