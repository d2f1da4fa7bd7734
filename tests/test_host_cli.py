"""Host side of the drop-in: libfaldoi_host.so (I/O + main()'s preprocessing) and the
`global_faldoi` executable's argv contract.  CPU tests do not need a GPU; the end-to-end
CLI runs are marked gpu."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_case

PKG = os.path.join(ROOT, "faldoi-ipol_b200")
BIN = os.path.join(PKG, "bin", "global_faldoi")
HOST_SO = os.path.join(PKG, "libfaldoi_host.so")


@pytest.fixture(scope="module")
def H():
    if not os.path.exists(HOST_SO) or not os.path.exists(BIN):
        subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "host")])
    L = C.CDLL(HOST_SO)
    L.faldoi_host_last_error.restype = C.c_char_p
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_host_exports(H):
    txt = open(os.path.join(ROOT, "include", "faldoi_host.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    syms = sorted(set(re.findall(r"\b(faldoi_host_[a-zA-Z0-9_]+)\s*\(", txt)))
    assert len(syms) == 7
    for s in syms:
        assert hasattr(H, s), s


@pytest.mark.parametrize("case", ["crop_a", "crop_b"])
def test_host_preprocess_matches_reference(H, case):
    g = load_case(case)
    rgb = [np.ascontiguousarray(g[k].astype(np.float32)) for k in ("rgb_i0", "rgb_i1", "rgb_im1")]
    _, h, w = rgb[0].shape
    out = [np.empty((h, w), np.float32) for _ in range(3)]
    assert H.faldoi_host_preprocess(_p(rgb[0]), _p(rgb[1]), _p(rgb[2]), 3, w, h, *[_p(o) for o in out]) == 0
    assert np.array_equal(out[0], g["I0n"]) and np.array_equal(out[1], g["I1n"]) and np.array_equal(out[2], g["Im1n"])
    lab = np.empty_like(rgb[0])
    assert H.faldoi_host_image_to_lab(_p(rgb[0]), w, h, _p(lab)) == 0
    assert np.array_equal(lab, g["lab"])


def _read(H, path):
    d, w, h, pd = C.POINTER(C.c_float)(), C.c_int(), C.c_int(), C.c_int()
    rc = H.faldoi_host_read_image(path.encode(), C.byref(d), C.byref(w), C.byref(h), C.byref(pd))
    if rc != 0:
        raise RuntimeError(H.faldoi_host_last_error().decode())
    a = np.ctypeslib.as_array(d, shape=(pd.value, h.value, w.value)).copy()
    H.faldoi_host_free(d)
    return a


def test_image_io_roundtrips(H, tmp_path, po):
    from PIL import Image
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    # PNG written by PIL (all filter types appear with optimize) -> our decoder
    for name, arr in (("rgb.png", rgb), ("gray.png", rgb[:, :, 0]), ("rgba.png", np.dstack([rgb, rgb[:, :, :1]]))):
        Image.fromarray(arr).save(tmp_path / name, optimize=True)
        a = _read(H, str(tmp_path / name))
        want = arr if arr.ndim == 3 else arr[:, :, None]
        assert np.array_equal(a, want.transpose(2, 0, 1).astype(np.float32)), name
    Image.fromarray((rgb[:, :, 0].astype(np.uint16) * 257)).save(tmp_path / "g16.png")
    assert np.array_equal(_read(H, str(tmp_path / "g16.png"))[0], rgb[:, :, 0].astype(np.float32) * 257)
    po.write_ppm(str(tmp_path / "a.ppm"), rgb)
    assert np.array_equal(_read(H, str(tmp_path / "a.ppm")), rgb.transpose(2, 0, 1).astype(np.float32))
    # .flo both ways
    u = rng.standard_normal((2, 37, 53)).astype(np.float32)
    assert H.faldoi_host_write_flo(str(tmp_path / "u.flo").encode(), _p(u[0]), _p(u[1]), 53, 37) == 0
    assert np.array_equal(po.read_flo(str(tmp_path / "u.flo")), u)
    assert np.array_equal(_read(H, str(tmp_path / "u.flo")), u)
    # occlusion mask PNG: values {0,1} in an 8-bit gray PNG, as iio_save_image_int writes it
    m = (rng.random((37, 53)) > 0.5).astype(np.int32)
    assert H.faldoi_host_write_png_gray8(str(tmp_path / "m.png").encode(), _p(m), 53, 37) == 0
    im = Image.open(tmp_path / "m.png")
    assert im.mode == "L" and np.array_equal(np.asarray(im), m.astype(np.uint8))
    with pytest.raises(RuntimeError):
        _read(H, str(tmp_path / "missing.png"))


def _write_case(tmp_path, g, po, frames=4):
    names = []
    for k, key in enumerate(("rgb_i0", "rgb_i1", "rgb_im1", "rgb_i1")[:frames]):
        p = str(tmp_path / ("f%d.ppm" % k))
        po.write_ppm(p, np.ascontiguousarray(g[key].transpose(1, 2, 0)))
        names.append(p)
    (tmp_path / "ims.txt").write_text("\n".join(names) + "\n")
    po.write_flo(str(tmp_path / "in.flo"), g["u0"])
    return str(tmp_path / "ims.txt"), str(tmp_path / "in.flo")


def test_cli_argv_contract(H, tmp_path, po):
    # wrong number of positionals -> usage on stderr, EXIT_FAILURE (src/global_faldoi.cpp:1865-1876)
    r = subprocess.run([BIN, "a", "b", "-m", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr and "today is:" in r.stderr
    g = load_case("crop_b")
    ims, flo = _write_case(tmp_path, g, po)
    # flow of the wrong size -> message + non-zero (:1961-1962)
    po.write_flo(str(tmp_path / "bad.flo"), g["u0"][:, :-1])
    r = subprocess.run([BIN, ims, str(tmp_path / "bad.flo"), str(tmp_path / "o.flo")], capture_output=True, text=True)
    assert r.returncode != 0 and "input flow field size mismatch" in r.stderr
    # -p <missing file>: the reference aborts in std::stof; we fail with a message
    r = subprocess.run([BIN, ims, flo, str(tmp_path / "o.flo"), "-p", "0"], capture_output=True, text=True)
    assert r.returncode != 0 and "parameter file" in r.stderr
    # options may appear anywhere; on a box without a GPU the solve must FAIL, not fall back
    r = subprocess.run([BIN, "-w", "1", ims, flo, "-m", "0", str(tmp_path / "o.flo")], capture_output=True, text=True)
    fb_has_gpu = os.path.exists("/dev/nvidia0")
    if not fb_has_gpu:
        assert r.returncode != 0 and "GPU solver failed" in r.stderr and not os.path.exists(tmp_path / "o.flo")


@pytest.mark.gpu
@pytest.mark.parametrize("method,warps,iters", [(0, 3, 400), (4, 1, 400), (2, 1, 400), (8, 1, 12)])
def test_cli_end_to_end(H, tmp_path, po, method, warps, iters):
    g = load_case("crop_b")
    ims, flo = _write_case(tmp_path, g, po)
    out = str(tmp_path / "out.flo")
    cmd = [BIN, ims, flo, out]
    if method == 8:
        from PIL import Image
        Image.fromarray(g["chi0"].astype(np.uint8)).save(tmp_path / "occ_in.png")
        cmd += [str(tmp_path / "occ_in.png"), str(tmp_path / "occ_out.png")]
    cmd += ["-m", str(method), "-w", str(warps), "-glb_iters", str(iters), "-verbose", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    u = po.read_flo(out)
    key = "m%d_w%d" % (method, warps) + ("_i%d" % iters if method == 8 else "")
    if method in (0, 4, 8):
        assert np.array_equal(u, g["u_" + key])
    else:
        # device preprocessing: Lab through the device's double pow (last-bit differences possible)
        d = np.abs(u - g["u_" + key])
        assert d.mean() <= 1e-3 and d.max() <= 1e-2
        # the reference's own preprocessing on the host: bit-identical flow
        r2 = subprocess.run(cmd + ["-host_preproc", "1"], capture_output=True, text=True)
        assert r2.returncode == 0, r2.stderr
        assert np.array_equal(po.read_flo(out), g["u_" + key])
    log = r.stderr if method in (0, 4) else r.stdout
    assert len(re.findall(r"Warping: \d+, ?Iter: \d+ Error: ", log)) == warps
    if method == 0:
        assert "(tvl2OF) All tasks took" in r.stdout
    if method == 8:
        from PIL import Image
        m = np.asarray(Image.open(tmp_path / "occ_out.png"))
        assert np.array_equal(m.astype(np.float32), g["chi_" + key])


@pytest.mark.gpu
def test_cli_two_frames_falls_back_from_occ(H, tmp_path, po):
    """m=8 with a 2-line ims.txt -> notice + TV-l2 coupled (src/global_faldoi.cpp:1965-1973)."""
    from PIL import Image
    g = load_case("crop_b")
    ims, flo = _write_case(tmp_path, g, po, frames=2)
    Image.fromarray(g["chi0"].astype(np.uint8)).save(tmp_path / "occ_in.png")
    out = str(tmp_path / "out.flo")
    r = subprocess.run([BIN, ims, flo, out, str(tmp_path / "occ_in.png"), str(tmp_path / "occ_out.png"), "-m", "8", "-w", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "method is changed to TV-l2 coupled" in r.stderr
    assert os.path.exists(out)


@pytest.mark.gpu
def test_cli_fullsize_real_pair_is_bit_identical_to_reference(H, tmp_path, po):
    """The drop-in claim end to end: same PNG frames, same local_faldoi flow, our executable (own PNG
    decoder, device preprocessing, GPU solver) writes the same .flo as the reference executable did."""
    D = os.path.join(ROOT, "oracle", "_ref", "data", "clean_easy")
    if not os.path.exists(os.path.join(D, "var_m0.flo")):
        pytest.skip("oracle/_ref/data/clean_easy not present")
    names = [os.path.join(D, "frame_%04d.png" % k) for k in (2, 3, 1, 4)]
    (tmp_path / "ims.txt").write_text("\n".join(names) + "\n")
    out = str(tmp_path / "out.flo")
    import time
    t = time.time()
    r = subprocess.run([BIN, str(tmp_path / "ims.txt"), os.path.join(D, "rg.flo"), out, "-m", "0", "-w", "5", "-verbose", "1"],
                       capture_output=True, text=True)
    wall = time.time() - t
    assert r.returncode == 0, r.stderr
    iters = [int(x) for x in re.findall(r"Warping: \d+,Iter: (\d+)", r.stderr)]
    assert iters == [400, 152, 100, 82, 136]
    assert open(out, "rb").read() == open(os.path.join(D, "var_m0.flo"), "rb").read()
    print("CLI wall %.2f s; %s" % (wall, [l for l in r.stdout.splitlines() if "All tasks" in l]))


@pytest.mark.gpu
@pytest.mark.parametrize("method", [4, 7, 8])
def test_cli_fullsize_other_models_byte_identical(H, tmp_path, po, method):
    """Same claim for TV-CSAD, NLTV-CSAD(-W) and TVL2-OCC: the .flo (and the occlusion PNG's pixels) our
    executable writes for the Sintel pair equal what the reference executable wrote (oracle/run_full_refs.sh)."""
    D = os.path.join(ROOT, "oracle", "_ref", "data", "clean_easy")
    ref = os.path.join(D, "var_m%d.flo" % method)
    flo = os.path.join(D, "rg_m8.flo" if method == 8 else "rg.flo")
    if not os.path.exists(ref) or not os.path.exists(flo):
        pytest.skip("full-size reference outputs not present")
    names = [os.path.join(D, "frame_%04d.png" % k) for k in (2, 3, 1, 4)]
    (tmp_path / "ims.txt").write_text("\n".join(names) + "\n")
    out = str(tmp_path / "out.flo")
    cmd = [BIN, str(tmp_path / "ims.txt"), flo, out]
    if method == 8:
        cmd += [os.path.join(D, "rg_occ.png"), str(tmp_path / "occ.png")]
    r = subprocess.run(cmd + ["-m", str(method), "-w", "5", "-glb_iters", "400"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(ref, "rb").read()
    if method == 8:
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(tmp_path / "occ.png")), np.asarray(Image.open(os.path.join(D, "var_m8_occ.png"))))


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 4])
def test_cli_fullsize_second_sequence_byte_identical(H, tmp_path, po, method):
    """A second Sintel pair (final/hard: motion blur, large displacements; oracle/run_full_refs_seq.sh): the CLI's
    .flo equals the reference executable's byte for byte, iteration counts and EPE as recorded in the golden.
    TV-CSAD is compared with a single-threaded reference run of one warp: with several OpenMP threads the
    reference's unsynchronised error sum (tvcsad_getP) loses most of its terms and ends the loop early."""
    D = os.path.join(ROOT, "oracle", "_ref", "data", "final_hard")
    ref = os.path.join(D, "var_m0.flo" if method == 0 else "var_m4_w1_t1.flo")
    gfile = os.path.join(ROOT, "tests", "golden", "fullsize_final_hard_m%d.npz" % method)
    if not os.path.exists(ref) or not os.path.exists(gfile):
        pytest.skip("final/hard reference outputs not present")
    g = dict(np.load(gfile))
    warps = 5 if method == 0 else 1
    names = [os.path.join(D, "frame_%04d.png" % k) for k in (2, 3, 1, 4)]
    (tmp_path / "ims.txt").write_text("\n".join(names) + "\n")
    out = str(tmp_path / "out.flo")
    r = subprocess.run([BIN, str(tmp_path / "ims.txt"), os.path.join(D, "rg.flo"), out, "-m", str(method), "-w", str(warps), "-verbose", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    iters = [int(x) for x in re.findall(r"Warping: \d+,Iter: (\d+)", r.stderr)]
    assert iters == list(g["iters"])
    assert open(out, "rb").read() == open(ref, "rb").read()
    u = po.read_flo(out)
    assert np.array_equal(u[:, ::8, ::8], g["u_sub"])
    gt = po.read_flo(os.path.join(D, "gt_frame_0002.flo"))
    assert round(float(np.sqrt(((u - gt) ** 2).sum(0)).mean()), 3) == round(float(g["epe_out"]), 3)


@pytest.mark.gpu
@pytest.mark.parametrize("seq", ["clean_medium", "clean_hard", "final_easy", "final_medium"])
def test_cli_fullsize_tvl2_remaining_sequences(H, tmp_path, po, seq):
    """TVL2 on the other example sequences of the reference: byte-identical .flo, same
    per-warp iteration counts, EPE equal to 3 decimals."""
    D = os.path.join(ROOT, "oracle", "_ref", "data", seq)
    ref = os.path.join(D, "var_m0.flo")
    gfile = os.path.join(ROOT, "tests", "golden", "fullsize_%s_m0.npz" % seq)
    if not os.path.exists(ref) or not os.path.exists(gfile):
        pytest.skip("%s reference outputs not present" % seq)
    g = dict(np.load(gfile))
    names = [os.path.join(D, "frame_%04d.png" % k) for k in (2, 3, 1, 4)]
    (tmp_path / "ims.txt").write_text("\n".join(names) + "\n")
    out = str(tmp_path / "out.flo")
    r = subprocess.run([BIN, str(tmp_path / "ims.txt"), os.path.join(D, "rg.flo"), out, "-m", "0", "-w", "5", "-verbose", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert [int(x) for x in re.findall(r"Warping: \d+,Iter: (\d+)", r.stderr)] == list(g["iters"])
    assert open(out, "rb").read() == open(ref, "rb").read()
    gt = po.read_flo(os.path.join(D, "gt_frame_0002.flo"))
    u = po.read_flo(out)
    assert round(float(np.sqrt(((u - gt) ** 2).sum(0)).mean()), 3) == round(float(g["epe_out"]), 3)


@pytest.mark.gpu
def test_cli_sequence_mode(H, tmp_path, po):
    """-seq jobs.txt: several pairs in one process (one CUDA start-up), each result as in a single call."""
    g = load_case("crop_b")
    ims, flo = _write_case(tmp_path, g, po)
    outs = [str(tmp_path / ("o%d.flo" % k)) for k in range(3)]
    (tmp_path / "jobs.txt").write_text("".join("%s %s %s\n" % (ims, flo, o) for o in outs))
    r = subprocess.run([BIN, "-seq", str(tmp_path / "jobs.txt"), "-m", "0", "-w", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "sequence: 3 pairs done" in r.stderr
    for o in outs:
        assert np.array_equal(po.read_flo(o), g["u_m0_w3"])


@pytest.mark.gpu
def test_cli_sequence_batched_mixed_sizes(H, tmp_path, po):
    """-seq fills the slots of a batched handle with consecutive jobs of equal size (-batch 2 here: batches of 2, 1, 2, 1
    as the size changes), through pinned staging; every result equals the single-call result.  -devices all shards
    the batches over every visible GPU."""
    ga, gb = load_case("crop_a"), load_case("crop_b")
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    ims_a, flo_a = _write_case(tmp_path / "a", ga, po)
    ims_b, flo_b = _write_case(tmp_path / "b", gb, po)
    order = "aaabbb"
    outs = [str(tmp_path / ("o%d.flo" % k)) for k in range(len(order))]
    (tmp_path / "jobs.txt").write_text("".join("%s %s %s\n" % ((ims_a, flo_a, o) if c == "a" else (ims_b, flo_b, o))
                                               for c, o in zip(order, outs)))
    for extra in ([], ["-devices", "all"]):
        r = subprocess.run([BIN, "-seq", str(tmp_path / "jobs.txt"), "-m", "0", "-w", "3", "-batch", "2", "-verbose", "1"] + extra,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "sequence: 6 pairs done" in r.stderr
        ua, _, ia, _ = po.o_global_solve(0, ga["I0n"], ga["I1n"], None, None, ga["u0"], warps=3)
        for c, o in zip(order, outs):
            assert np.array_equal(po.read_flo(o), ua if c == "a" else gb["u_m0_w3"])
            os.remove(o)
        assert r.stderr.count("Warping: 0,Iter:") == 6


def test_cli_unknown_method_writes_input_flow(H, tmp_path, po):
    """Method ids outside 0..8: the reference's dispatch matches nothing and main() saves the input flow unchanged."""
    g = load_case("crop_b")
    ims, flo = _write_case(tmp_path, g, po)
    out = str(tmp_path / "out.flo")
    r = subprocess.run([BIN, ims, flo, out, "-m", "-1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(flo, "rb").read()


@pytest.mark.gpu
def test_cli_sequence_pipeline_occ_and_errors(H, tmp_path, po):
    """-seq with method 8 (flow + occlusion PNG per job, written by the background writer) and a job whose
    input is missing: the run stops with a non-zero code, the jobs before it are complete on disk."""
    from PIL import Image
    g = load_case("crop_b")
    ims, flo = _write_case(tmp_path, g, po)
    Image.fromarray(g["chi0"].astype(np.uint8)).save(tmp_path / "occ_in.png")
    occ_in = str(tmp_path / "occ_in.png")
    jobs = ["%s %s %s %s %s" % (ims, flo, tmp_path / ("o%d.flo" % k), occ_in, tmp_path / ("m%d.png" % k)) for k in range(4)]
    (tmp_path / "jobs.txt").write_text("\n".join(jobs) + "\n")
    opts = ["-m", "8", "-w", "1", "-glb_iters", "12"]
    r = subprocess.run([BIN, "-seq", str(tmp_path / "jobs.txt")] + opts, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "sequence: 4 pairs done" in r.stderr
    for k in range(4):
        assert np.array_equal(po.read_flo(str(tmp_path / ("o%d.flo" % k))), g["u_m8_w1_i12"])
        assert np.array_equal(np.asarray(Image.open(tmp_path / ("m%d.png" % k))).astype(np.float32), g["chi_m8_w1_i12"])
    # job 2 of 3 points at a flow file that does not exist
    bad = [jobs[0].replace("o0.flo", "p0.flo").replace("m0.png", "q0.png"),
           "%s %s %s %s %s" % (ims, tmp_path / "nope.flo", tmp_path / "p1.flo", occ_in, tmp_path / "q1.png"),
           jobs[2].replace("o2.flo", "p2.flo").replace("m2.png", "q2.png")]
    (tmp_path / "bad.txt").write_text("\n".join(bad) + "\n")
    r = subprocess.run([BIN, "-seq", str(tmp_path / "bad.txt")] + opts, capture_output=True, text=True)
    assert r.returncode != 0 and "ERROR" in r.stderr
    assert np.array_equal(po.read_flo(str(tmp_path / "p0.flo")), g["u_m8_w1_i12"])
    assert not os.path.exists(tmp_path / "p1.flo") and not os.path.exists(tmp_path / "p2.flo")


def test_cli_sequence_pipeline_without_gpu(H, tmp_path, po):
    """The host side of -seq (decode pool, staging slots, batching by size, writer pool, stop at the first failing
    job) exercised without a GPU: method id -1 makes every job a pass-through (the reference writes the input flow
    back for ids outside 0..8), so results are known and no solver is needed."""
    ga, gb = load_case("crop_a"), load_case("crop_b")
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    ims_a, flo_a = _write_case(tmp_path / "a", ga, po)
    ims_b, flo_b = _write_case(tmp_path / "b", gb, po)
    order = "aabbbabab" * 3
    outs = [str(tmp_path / ("o%02d.flo" % k)) for k in range(len(order))]
    (tmp_path / "jobs.txt").write_text("".join("%s %s %s\n" % ((ims_a, flo_a, o) if c == "a" else (ims_b, flo_b, o))
                                               for c, o in zip(order, outs)))
    r = subprocess.run([BIN, "-seq", str(tmp_path / "jobs.txt"), "-m", "-1", "-batch", "3", "-seq_stats", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "sequence: %d pairs done" % len(order) in r.stderr and "pipeline: %d pairs" % len(order) in r.stderr
    for c, o in zip(order, outs):
        assert open(o, "rb").read() == open(flo_a if c == "a" else flo_b, "rb").read()
    # a job in the middle whose flow file is missing: non-zero exit, earlier jobs complete, later jobs not started
    lines = (tmp_path / "jobs.txt").read_text().splitlines()
    for o in outs:
        os.remove(o)
    lines[7] = "%s %s %s" % (ims_a, tmp_path / "nope.flo", outs[7])
    (tmp_path / "bad.txt").write_text("\n".join(lines) + "\n")
    r = subprocess.run([BIN, "-seq", str(tmp_path / "bad.txt"), "-m", "-1", "-batch", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and "ERROR" in r.stderr
    assert all(os.path.exists(o) for o in outs[:7]) and not any(os.path.exists(o) for o in outs[7:])
    # a job line with the wrong number of file names is refused before anything runs
    (tmp_path / "short.txt").write_text("%s %s\n" % (ims_a, flo_a))
    r = subprocess.run([BIN, "-seq", str(tmp_path / "short.txt"), "-m", "-1"], capture_output=True, text=True)
    assert r.returncode != 0 and "needs 3 or 5 file names" in r.stderr


def test_png_header_is_validated(H, tmp_path):
    """Untrusted PNG headers: bit depths the specification does not allow for the colour type and absurd sizes are
    refused before any allocation or shift uses them."""
    import struct
    import zlib

    def png(w, h, depth, ctype, payload=b"\0" * 16):
        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
        return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(payload)) + chunk(b"IEND", b""))

    for name, data, msg in (("depth0.png", png(4, 4, 0, 0), "bit depth"), ("depth32.png", png(4, 4, 32, 2), "bit depth"),
                            ("rgb4.png", png(4, 4, 4, 2), "bit depth"), ("huge.png", png(0x7fffffff, 0x7fffffff, 8, 0), "too large")):
        path = tmp_path / name
        path.write_bytes(data)
        with pytest.raises(RuntimeError) as e:
            _read(H, str(path))
        assert msg in str(e.value)
    # a valid 2x1 gray image still decodes
    ok = tmp_path / "ok.png"
    ok.write_bytes(png(2, 1, 8, 0, b"\0\x07\x09"))
    a = _read(H, str(ok))
    assert a.shape == (1, 1, 2) and a[0, 0, 0] == 7 and a[0, 0, 1] == 9


def test_write_errors_are_reported(H, tmp_path, po):
    """A full disk (/dev/full) while writing the flow must end in an error, not in a truncated file behind a successful exit."""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full here")
    u = np.zeros((2, 64, 64), np.float32)
    assert H.faldoi_host_write_flo(b"/dev/full", _p(u[0]), _p(u[1]), 64, 64) != 0
    assert b"writing" in H.faldoi_host_last_error()
    g = load_case("crop_b")
    ims, flo = _write_case(tmp_path, g, po)
    r = subprocess.run([BIN, ims, flo, "/dev/full", "-m", "-1"], capture_output=True, text=True)
    assert r.returncode != 0 and "ERROR" in r.stderr
