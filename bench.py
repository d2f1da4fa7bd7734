#!/usr/bin/env python
"""bench.py -- global_faldoi throughput on B200 (BASELINE.json metric: Mpix*iter/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU implementation, rank 0 only

A "step" = one global_faldoi pass (all warps, all iterations) over one batch of synthetic
1024x436 pairs per GPU (BASELINE.json configs[4]: the 64-pair sequence sharded by pair; default
64 pairs resident per GPU, i.e. the whole named sequence on every GPU).  Pairs are independent
-> no data-path collective, weak scaling.  Work is counted as pixels x iterations actually run
(the TVL2 loop exits early per pair, src/global_faldoi.cpp:684), so launches that find a pair
already converged cost time but add no work.

  value  : inputs resident in HBM (device->device staging of the 4 input planes + solve), CUDA events
           on the solver's stream, max over ranks.
  e2e    : the same through the C-ABI with pinned HOST buffers: H2D of I0,I1,u per pair, solve, D2H of
           the flow, wall clock between device syncs, max over ranks.
  roofline: the iteration kernel (tv_tile2_kernel, two iterations per launch): algorithmic 80 B/px/iter
            x pixel-iterations run / time inside the iteration launches (events around
           every warp's iteration loop, faldoi_solver_last_iter_ms) vs the measured HBM copy bandwidth;
           `whole_step` is the same over the whole step (bicubic warps, constants, staging included).
           `traffic` = DRAM bytes of ONE launch (ncu capture under profiles/, scaled to this batch).
  cpu_baseline: the UNMODIFIED reference tvl2OF (oracle/_ref, OpenMP, all host cores) on ONE pair of
           the same workload (about 5-10 s); falls back to the C port (oracle/) if the reference
           build did not travel.
  methods (N=1): the other configurations BASELINE.json names, measured the same way after the headline:
           TV-CSAD (method 4), NLTV-CSAD-W (7), TVL2-OCC (8) on 1024x436 pairs and TVL2 on a 3840x2160 pair,
           each with value / e2e / roofline and a bit-equality check of the GPU flow against the reference
           build on a bounded sample (stated per entry).
  stripes (N>1): after the pair-sharded headline rank 0 cuts ONE 3840x2160 pair into N row stripes over the
           N GPUs (halo rows exchanged over NVLink by the iteration kernel) and reports value, efficiency
           against its own single-GPU solve and whether the flows are bit-identical.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = {0: 80, 1: 80, 4: 84, 5: 84, 2: 532, 3: 532, 6: 536, 7: 536, 8: 2370}  # SURVEY.md 8(d), per px*iter
KERNEL = {0: "tv_tile2_kernel", 1: "tv_tile2_kernel", 4: "tv_csad2_kernel", 5: "tv_csad2_kernel", 2: "nltv_tile_kernel",
          3: "nltv_tile_kernel", 6: "nltv_tile_kernel<CSAD>", 7: "nltv_tile_kernel<CSAD>", 8: "occ_xi_rows_kernel + occ_chi_rows_kernel"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=64,
                    help="pairs resident per GPU; 64 = the whole named 64-pair sequence on every GPU (weak scaling)")
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--height", type=int, default=436)
    ap.add_argument("--method", type=int, default=0)
    ap.add_argument("--warps", type=int, default=5)
    ap.add_argument("--glb-iters", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the `methods` (N=1) / `stripes` (N>1) keys")
    ap.add_argument("--mode", default="pairs", choices=["pairs", "stripes"],
                    help="pairs: independent pairs per GPU (default, the driver's line). stripes: ONE frame pair cut into "
                         "row stripes over --gpus GPUs with halo exchange over NVLink (3840x2160 by default); one host "
                         "process drives all GPUs (under torchrun rank 0 does, the other ranks exit)")
    return ap.parse_args()


def pin_omp_threads():
    """The reference is OpenMP code and takes its team size from OMP_NUM_THREADS; torch.distributed.run exports
    OMP_NUM_THREADS=1 to every worker, which would time a single-threaded reference.  Pin it to the host's
    core count (environment for a libgomp not yet initialised, omp_set_num_threads for one that is)."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL).omp_set_num_threads(n)
    except OSError:
        pass
    return n


# ----------------------------------------------------------------------------- synthetic workload
def make_pairs(npairs, w, h, seed0, device):
    """Seeded synthetic pairs: band-limited texture, smooth true flow, I1 / I-1 = I0
    warped along +/- flow, init flow = truth + smooth noise (stands in for local_faldoi).  Generated
    with torch on `device` (plumbing), returned as float32 tensors [npairs, ...] on that device."""
    import torch
    import torch.nn.functional as F

    def blur(x, sigma):
        k = int(3 * sigma) | 1
        t = torch.arange(-k, k + 1, device=x.device, dtype=torch.float32)
        ker = torch.exp(-0.5 * (t / sigma) ** 2)
        ker = ker / ker.sum()
        x = F.conv2d(F.pad(x, (k, k, 0, 0), mode="reflect"), ker.view(1, 1, 1, -1))
        return F.conv2d(F.pad(x, (0, 0, k, k), mode="reflect"), ker.view(1, 1, -1, 1))

    out = {k: [] for k in ("I0", "I1", "Im1", "u0", "rgb")}
    yy, xx = torch.meshgrid(torch.arange(h, device=device, dtype=torch.float32),
                            torch.arange(w, device=device, dtype=torch.float32), indexing="ij")
    for k in range(npairs):
        g = torch.Generator(device=device)
        g.manual_seed(seed0 + k)
        tex = sum(blur(torch.randn(1, 1, h, w, device=device, generator=g), 2.0 ** o) * (2.0 ** o) for o in range(5))
        tex = (tex - tex.min()) / (tex.max() - tex.min())
        ph = torch.rand(4, device=device, generator=g) * 6.28
        amp = 2.0 + 4.0 * torch.rand(1, device=device, generator=g)
        f1 = amp * 0.5 * (torch.sin(6.28 * xx / w + ph[0]) + 0.5 * torch.cos(6.28 * yy / h + ph[1]))
        f2 = amp * 0.5 * (torch.cos(9.42 * xx / w + ph[2]) - 0.5 * torch.sin(6.28 * yy / h + ph[3]))

        def warp(img, s):
            gx = (xx + s * f1) / (w - 1) * 2 - 1
            gy = (yy + s * f2) / (h - 1) * 2 - 1
            return F.grid_sample(img, torch.stack([gx, gy], -1)[None], mode="bicubic", padding_mode="border",
                                 align_corners=True)

        I0 = blur(tex, 0.9)
        I1 = blur(warp(tex, 1.0), 0.9)
        Im1 = blur(warp(tex, -1.0), 0.9)
        n1 = blur(torch.randn(1, 1, h, w, device=device, generator=g), 3.0)[0, 0] * 3.0
        n2 = blur(torch.randn(1, 1, h, w, device=device, generator=g), 3.0)[0, 0] * 3.0
        out["I0"].append(I0[0, 0].clamp(0, 1))
        out["I1"].append(I1[0, 0].clamp(0, 1))
        out["Im1"].append(Im1[0, 0].clamp(0, 1))
        out["u0"].append(torch.stack([f1 + 0.5 * n1, f2 + 0.5 * n2]))
        out["rgb"].append(torch.stack([tex[0, 0] * 255, tex[0, 0].roll(3, 1) * 200 + 20, tex[0, 0].roll(5, 0) * 180 + 40]))
    return {k: torch.stack(v).contiguous().float() for k, v in out.items()}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms (the checker / the baseline)
def pyoracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    return po


def cpu_solve_one(I0, I1, u0, warps, method, Im1=None, lab=None, chi=None, glb_iters=400, deterministic_port=False):
    """Times the reference's own CPU implementation on one pair.  Returns (seconds, kind, cores, u_out, chi_out).
    deterministic_port: use the C restatement even when the reference build is there (TV-CSAD: the reference's
    `err_D +=` inside an OpenMP loop is a race that feeds its exit test, so only a single-threaded reference run
    is reproducible -- the port sums in order and equals that run bit for bit, tests/test_oracle_vs_ref.py)."""
    po = pyoracle()
    cores = pin_omp_threads()
    if po.have_ref() and not deterministic_port:
        t = time.perf_counter()
        u, c = po.r_global_solve(method, I0, I1, Im1, lab, u0, chi, warps=warps, glb_iters=glb_iters)
        return time.perf_counter() - t, "reference", cores, u, c
    t = time.perf_counter()
    u, c, _, _ = po.o_global_solve(method, I0, I1, Im1, lab, u0, chi, warps=warps, glb_iters=glb_iters)
    return time.perf_counter() - t, "port", cores, u, c


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(key, npairs_scale):
    """DRAM bytes of one launch of the dominant kernel from the committed ncu capture (profiles/traffic.json:
    {key: {"bytes": per-launch bytes, "pairs": pairs resident in the capture, "file": csv}}), scaled to this batch."""
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(prof):
        return None
    e = json.load(open(prof)).get(str(key))
    if e is None:
        return None
    if isinstance(e, dict):
        return e["bytes"] * npairs_scale / float(e.get("pairs", 1))
    return e * npairs_scale / 16.0  # round-1 format: 16-pair capture


# ----------------------------------------------------------------------------- one measured configuration
class Workload:
    """`B` synthetic pairs of w x h for `method`, resident on `dev` and mirrored in pinned host memory."""

    def __init__(self, fb, torch, dev, method, w, h, B, warps, glb_iters, seed0):
        import numpy as np
        self.fb, self.torch, self.dev = fb, torch, dev
        self.method, self.w, self.h, self.B, self.warps, self.glb_iters = method, w, h, B, warps, glb_iters
        self.npix = w * h
        d = make_pairs(B, w, h, seed0, dev)
        self.need_lab = method in (2, 3, 6, 7)
        lab = None
        if self.need_lab:
            # Lab conversion is preprocessing (once, outside the timed region): the product's own host
            # library does it (libfaldoi_host.so, include/faldoi_host.h) -- not the oracle.
            import ctypes as C
            hostlib = C.CDLL(os.path.join(ROOT, "faldoi-ipol_b200", "libfaldoi_host.so"))
            rgb = np.ascontiguousarray(d["rgb"].cpu().numpy())
            lab_np = np.empty_like(rgb)
            for k in range(B):
                hostlib.faldoi_host_image_to_lab(rgb[k].ctypes.data_as(C.c_void_p), w, h, lab_np[k].ctypes.data_as(C.c_void_p))
            lab = torch.from_numpy(lab_np).to(dev)
        del d["rgb"]
        chi = torch.zeros(B, h, w, device=dev) if method == 8 else None
        self.d, self.lab, self.chi = d, lab, chi
        self.host = {k: d[k].cpu().pin_memory() for k in ("I0", "I1", "Im1", "u0")}
        self.host_lab = lab.cpu().pin_memory() if lab is not None else None
        self.host_chi = chi.cpu().pin_memory() if chi is not None else None
        self.host_out = torch.empty(B, 2, h, w).pin_memory()
        self.host_chi_out = torch.empty(B, h, w).pin_memory() if method == 8 else None
        self.solver = fb.Solver(w, h, method, B, device=dev.index)
        self.params = fb.default_params(method, glb_iters, warps)
        self.stream = torch.cuda.ExternalStream(self.solver.stream, device=dev)

    def _stage(self, src, lab_t, chi_t):
        def ptr(t, k):
            return t[k].data_ptr() if t is not None else 0
        m8 = self.method == 8
        for k in range(self.B):
            self.solver.upload_ptrs(k, ptr(src["I0"], k), ptr(src["I1"], k), ptr(src["u0"], k),
                                    Im1=ptr(src["Im1"], k) if m8 else 0, lab=ptr(lab_t, k), chi=ptr(chi_t, k))

    def step_resident(self):
        self._stage(self.d, self.lab, self.chi)
        self.solver.run(self.params)

    def step_e2e(self):
        self._stage(self.host, self.host_lab, self.host_chi)
        self.solver.run(self.params)
        for k in range(self.B):
            self.solver.download_ptr(k, self.host_out[k].data_ptr(),
                                     self.host_chi_out[k].data_ptr() if self.host_chi_out is not None else 0)

    def measure(self, steps, warmup, barrier):
        """-> dict(t_res, t_e2e, iter_ms, launches, units_step, iters_total).  `barrier` syncs the device (and the ranks)."""
        torch = self.torch
        for _ in range(max(warmup, 1)):  # also yields the per-pair iteration counts: deterministic, identical every step
            self.step_resident()
            self.solver.sync()
        iters_total = 0
        for k in range(self.B):
            _, _, log = self.solver.download(k)
            iters_total += sum(log.iters[:self.warps])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        iter_ms, launches = 0.0, 0
        for _ in range(steps):
            self.step_resident()
            self.solver.sync()  # needed to read last_iter_ms; the stream stays back-to-back within a step
            iter_ms += self.solver.last_iter_ms
            launches += self.solver.last_launches
        e1.record(self.stream)
        barrier()
        t_res = e0.elapsed_time(e1) / 1e3
        self.step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step_e2e()
        barrier()
        t_e2e = time.perf_counter() - t0
        return {"t_res": t_res, "t_e2e": t_e2e, "iter_ms": iter_ms, "launches": launches,
                "units_step": self.npix * iters_total, "iters_total": iters_total}

    def bytes_in(self):
        return self.B * (4 + (2 if self.method == 8 else 0) + (3 if self.need_lab else 0)) * self.npix * 4

    def bytes_out(self):
        return self.B * (2 + (1 if self.method == 8 else 0)) * self.npix * 4

    def close(self):
        self.solver.close()


def roofline_entry(method, units, steps, iter_ms, t_res, traffic):
    peak, peak_src = peak_hbm()
    alg = ALG_BYTES[method]
    ach = alg * units * steps / (iter_ms / 1e3) / 1e9
    whole = alg * units * steps / t_res / 1e9
    return {"bound": "hbm", "kernel": KERNEL[method], "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "frac": ach / peak, "alg_bytes_per_px_iter": alg, "traffic": traffic,
            "whole_step": {"achieved": whole, "frac": whole / peak,
                           "note": "same bytes over the whole device-timed step (bicubic warps, per-warp constants, staging copies included)"}}


def parity_sample(fb, torch, dev, method, w, h, warps, glb_iters, seed):
    """GPU (through the C ABI, host buffers) against the reference build on ONE bounded pair."""
    import numpy as np
    d = make_pairs(1, w, h, seed, dev)
    I0, I1, Im1, u0 = [d[k][0].cpu().numpy() for k in ("I0", "I1", "Im1", "u0")]
    lab = chi = None
    if method in (2, 3, 6, 7):
        import ctypes as C
        hostlib = C.CDLL(os.path.join(ROOT, "faldoi-ipol_b200", "libfaldoi_host.so"))
        rgb = np.ascontiguousarray(d["rgb"][0].cpu().numpy())
        lab = np.empty_like(rgb)
        hostlib.faldoi_host_image_to_lab(rgb.ctypes.data_as(C.c_void_p), w, h, lab.ctypes.data_as(C.c_void_p))
    if method == 8:
        chi = np.zeros((h, w), np.float32)
    ug, cg, its, _ = fb.global_solve(method, I0, I1, u0, Im1=Im1 if method == 8 else None, lab=lab, chi=chi, warps=warps,
                                     glb_iters=glb_iters, device=dev.index)
    t, kind, cores, uc, cc = cpu_solve_one(I0, I1, u0, warps, method, Im1=Im1 if method == 8 else None, lab=lab, chi=chi,
                                           glb_iters=glb_iters, deterministic_port=method in (4, 5))
    same = bool(np.array_equal(ug, uc)) and (cg is None or bool(np.array_equal(cg, cc)))
    return same, {"checker": kind, "cores": cores, "cpu_s": round(t, 2), "iters": its,
                  "sample": "1 synthetic %dx%d pair (seed %d), %d warp(s)%s" % (w, h, seed, warps, ", %d outer iterations" % glb_iters if method == 8 else ""),
                  "cpu_value": w * h * sum(its) / t / 1e6, "max_abs_diff": float(np.abs(ug - uc).max())}


def extra_methods(fb, torch, dev, a):
    """BASELINE.json configs[1..4] on one GPU, after the headline (rank 0, N=1)."""
    def nobarrier():
        torch.cuda.synchronize()

    plans = [  # key, method, w, h, pairs, warps, glb_iters, steps, (parity w, h, warps, glb_iters)
        ("tvcsad", 4, 1024, 436, 16, 5, 400, 3, (300, 131, 1, 400)),
        ("nltvcsad_w", 7, 1024, 436, 16, 5, 400, 2, (300, 131, 1, 400)),
        ("tvl2_occ", 8, 1024, 436, 16, 5, 100, 2, (300, 131, 1, 12)),
        ("tvl2_3840x2160", 0, 3840, 2160, 2, 5, 400, 3, (3840, 2160, 1, 400)),
    ]
    out = {}
    for key, method, w, h, B, warps, glb, steps, par in plans:
        wl = Workload(fb, torch, dev, method, w, h, B, warps, glb, 3000 + 100 * method + (7 if w > 2000 else 0))
        m = wl.measure(steps, 3, nobarrier)
        units = m["units_step"]
        e = {"method": method, "pairs": B, "width": w, "height": h, "warps": warps,
             "max_iters": glb if method == 8 else 400, "steps": steps, "warmup": 3,
             "unit": "Mpix*outer-iter/s" if method == 8 else "Mpix*iter/s",
             "value": units * steps / m["t_res"] / 1e6, "ms": 1e3 * m["t_res"] / steps,
             "e2e": {"value": units * steps / m["t_e2e"] / 1e6, "h2d_bytes_per_step": wl.bytes_in(), "d2h_bytes_per_step": wl.bytes_out()},
             "pairs_per_s": B * steps / m["t_res"], "iters_per_pair": m["iters_total"] / B, "gpu_launches": int(m["launches"]),
             "roofline": roofline_entry(method, units, steps, m["iter_ms"], m["t_res"], traffic_for(key, B))}
        wl.close()
        del wl
        torch.cuda.empty_cache()
        same, info = parity_sample(fb, torch, dev, method, par[0], par[1], par[2], par[3], 4000 + method)
        e["bit_identical_to_reference"] = same
        e["parity"] = info
        out[key] = e
    return out


def stripes_section(fb, torch, a, ngpu):
    """ONE 3840x2160 pair in `ngpu` row stripes against its single-GPU solve (rank 0 drives every GPU)."""
    import numpy as np
    w, h = 3840, 2160
    npix = w * h
    dev0 = torch.device("cuda", 0)
    d = make_pairs(1, w, h, 2000, dev0)
    host = {k: d[k][0].cpu().pin_memory() for k in ("I0", "I1", "u0")}
    del d
    torch.cuda.empty_cache()
    params = fb.default_params(0, 400, a.warps)
    steps, warm = 3, 2
    res = {}
    flows = {}
    for n in (1, ngpu):
        g = fb.Stripes(w, h, list(range(n)))
        out = torch.empty(2, h, w).pin_memory()
        for _ in range(warm):
            g.upload_ptrs(host["I0"].data_ptr(), host["I1"].data_ptr(), host["u0"].data_ptr())
            g.run(params)
        u, log = g.download()
        iters = sum(log.iters[:a.warps])
        ms, launches = 0.0, 0
        for _ in range(steps):
            g.upload_ptrs(host["I0"].data_ptr(), host["I1"].data_ptr(), host["u0"].data_ptr())
            g.run(params)
            g.download_ptr(out.data_ptr())
            ms += g.last_run_ms
            launches += g.last_launches
        flows[n] = (u, list(log.iters[:a.warps]))
        res[n] = {"value": npix * iters * steps / (ms / 1e3) / 1e6, "ms": ms / steps, "iters": iters, "launches": int(launches)}
        g.close()
    peak, _ = peak_hbm()
    same = bool(np.array_equal(flows[1][0], flows[ngpu][0])) and flows[1][1] == flows[ngpu][1]
    return {"workload": "synthetic 3840x2160 pair (seed 2000), TVL2, %d warps x <=400 iters, one row stripe per GPU, halo rows "
                        "stored into the neighbour GPU over NVLink by the iteration kernel" % a.warps,
            "n_gpus": ngpu, "scaling": "strong", "unit": "Mpix*iter/s", "steps": steps, "warmup": warm,
            "value": res[ngpu]["value"], "ms": res[ngpu]["ms"], "value_1gpu": res[1]["value"], "ms_1gpu": res[1]["ms"],
            "efficiency_vs_1gpu": res[ngpu]["value"] / res[1]["value"] / ngpu,
            "roofline_frac": 80 * res[ngpu]["value"] * 1e6 / 1e9 / (peak * ngpu),
            "iters": res[ngpu]["iters"], "gpu_launches": res[ngpu]["launches"],
            "bit_identical_to_single_gpu": same}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    w, h, B, method = a.width, a.height, a.pairs_per_gpu, a.method
    npix = w * h
    cfg = {"workload": "synthetic %dx%d pairs, %d per GPU, sharded by pair; method %d (%s), %d warps x <=%d iters, tol 0.01"
                       % (w, h, B, method, "tvl2OF" if method < 2 else "m%d" % method, a.warps, a.glb_iters if method == 8 else 400),
           "pairs_per_gpu": B, "width": w, "height": h, "method": method, "warps": a.warps,
           "l2_policy": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (B * 24 * npix * 4 / 1e9)}
    base = {"metric": "global_faldoi Mpix*iter/s", "unit": "Mpix*iter/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg}

    import numpy as np

    if a.mode == "stripes":
        return main_stripes(a, rank, base)

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        cores = pin_omp_threads()  # before anything loads the reference's OpenMP runtime
        import torch
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        d = make_pairs(1, w, h, 1000, dev)
        I0, I1, u0 = [d[k][0].cpu().numpy() for k in ("I0", "I1", "u0")]
        po = pyoracle()
        # iteration counts of this pair (the unit of work): from the C port, untimed, before the reference is timed
        _, _, iters, _ = po.o_global_solve(method, I0, I1, None, None, u0, warps=a.warps) if method == 0 else (0, 0, [400] * a.warps, 0)
        times, kind = [], None
        for k in range(a.warmup + a.steps):
            t, kind, cores, _, _ = cpu_solve_one(I0, I1, u0, a.warps, method)
            if k >= a.warmup:
                times.append(t)
        tot = sum(times)
        val = npix * sum(iters) * a.steps / tot / 1e6
        line = dict(base, impl="reference", value=val, ms_per_step=1e3 * tot / a.steps, gpu_launches=0,
                    omp_num_threads=cores,
                    cpu_baseline={"value": val, "unit": "Mpix*iter/s", "cores": cores, "kind": kind,
                                  "sample": "1 pair (seed 1000) of the workload per step, %d iterations; OMP_NUM_THREADS pinned to %d "
                                            "(torchrun exports 1)" % (sum(iters), cores)},
                    e2e={"value": val, "unit": "Mpix*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fb = importlib.import_module("faldoi-ipol_b200")
    if method in (2, 3, 6, 7):
        cfg["nltv_arithmetic"] = "exact"
    wl = Workload(fb, torch, dev, method, w, h, B, a.warps, a.glb_iters, 1000 + rank * B)
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = wl.measure(a.steps, a.warmup, barrier)
    clocks = sampler.stop()
    units_step, iters_total = m["units_step"], m["iters_total"]

    tt = torch.tensor([m["t_res"], m["t_e2e"], m["iter_ms"]], device=dev, dtype=torch.float64)
    uu = torch.tensor([float(units_step), float(m["launches"])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(uu, op=dist.ReduceOp.SUM)
    t_res, t_e2e, _ = [float(x) for x in tt.cpu()]
    units_all, launches_all = [float(x) for x in uu.cpu()]
    if world > 1:
        # the data path is done: leave the process group so that the other ranks can exit (and free their GPUs)
        # while rank 0 goes on to the row-stripe section
        barrier()
        dist.destroy_process_group()

    if rank == 0:
        value = units_all * a.steps / t_res / 1e6
        e2e = units_all * a.steps / t_e2e / 1e6
        line = dict(base, value=value, ms_per_step=1e3 * t_res / a.steps,
                    e2e={"value": e2e, "unit": "Mpix*iter/s", "h2d_bytes_per_step": wl.bytes_in(),
                         "d2h_bytes_per_step": wl.bytes_out(), "pairs_per_s": world * B * a.steps / t_e2e},
                    pairs_per_s=world * B * a.steps / t_res, iters_per_pair=iters_total / B,
                    gpu_launches=int(launches_all),
                    # rank 0's own kernel: bytes it moved algorithmically / its time inside the iteration launches
                    roofline=roofline_entry(method, units_step, a.steps, m["iter_ms"], m["t_res"], traffic_for(method, B)),
                    clocks=clocks)
        if not a.no_cpu_baseline and world == 1 and method in (0, 1):
            k0 = 0
            I0, I1, u0 = [wl.host[k][k0].numpy() for k in ("I0", "I1", "u0")]
            _, _, log = wl.solver.download(k0)
            t, kind, cores, u_cpu, _ = cpu_solve_one(I0, I1, u0, a.warps, method)
            its = sum(log.iters[:a.warps])
            line["cpu_baseline"] = {"value": npix * its / t / 1e6, "unit": "Mpix*iter/s", "cores": cores, "kind": kind,
                                    "sample": "pair 0 of the workload, %d iterations, %.2f s; flow bit-identical to GPU: %s"
                                              % (its, t, bool(np.array_equal(u_cpu, wl.host_out[k0].numpy())))}
        wl.close()
        del wl
        torch.cuda.empty_cache()
        if not a.no_extras:
            try:
                if world == 1:
                    line["methods"] = extra_methods(fb, torch, dev, a)
                elif torch.cuda.device_count() >= world:
                    time.sleep(3.0)  # the other ranks are tearing down their contexts
                    line["stripes"] = stripes_section(fb, torch, a, world)
            except Exception as ex:  # the headline stands on its own; say what went wrong with the rest
                line["extras_error"] = "%s: %s" % (type(ex).__name__, ex)
        print(json.dumps(line))
    else:
        wl.close()


def main_stripes(a, rank, base):
    """BASELINE.json configs[4], second half: 3840x2160 pair, row stripes + 2-row halo exchange once per two-iteration launch."""
    import numpy as np
    import torch
    if rank != 0:
        return
    w, h = (a.width, a.height) if (a.width, a.height) != (1024, 436) else (3840, 2160)
    npix = w * h
    fb = importlib.import_module("faldoi-ipol_b200")
    ng = a.gpus
    if not torch.cuda.is_available() or torch.cuda.device_count() < ng:
        raise SystemExit("bench.py --mode stripes --gpus %d: not enough CUDA devices" % ng)
    d = make_pairs(1, w, h, 2000, torch.device("cuda", 0))
    host = {k: d[k][0].cpu().pin_memory() for k in ("I0", "I1", "u0")}
    out = torch.empty(2, h, w).pin_memory()
    params = fb.default_params(0, 400, a.warps)
    g = fb.Stripes(w, h, list(range(ng)))

    def step():
        g.upload_ptrs(host["I0"].data_ptr(), host["I1"].data_ptr(), host["u0"].data_ptr())
        g.run(params)
        g.download_ptr(out.data_ptr())

    for _ in range(max(a.warmup, 1)):
        step()
    _, log = g.download()
    iters = sum(log.iters[:a.warps])
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(a.steps):
        step()
        dev_ms += g.last_run_ms
        launches += g.last_launches
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop()
    units = npix * iters
    peak, peak_src = peak_hbm()
    ach = 80 * units * a.steps / (dev_ms / 1e3) / 1e9
    cfg = dict(base["config"], workload="synthetic %dx%d pair cut into %d row stripes (one per GPU), two halo rows per side exchanged once "
               "per two-iteration launch by NVLink peer stores from the iteration kernel; TVL2, %d warps x <=400 iters" % (w, h, ng, a.warps),
               width=w, height=h, pairs_per_gpu=None, l2_policy="state %.1f GB >> L2" % (23 * npix * 4 / 1e9))
    line = dict(base, config=cfg, scaling="strong", value=units * a.steps / (dev_ms / 1e3) / 1e6,
                ms_per_step=dev_ms / a.steps, iters_per_pair=iters, gpu_launches=int(launches),
                e2e={"value": units * a.steps / t_e2e / 1e6, "unit": "Mpix*iter/s", "h2d_bytes_per_step": (ng + 3) * npix * 4,
                     "d2h_bytes_per_step": 2 * npix * 4},
                roofline={"bound": "hbm", "kernel": "tv_tile2_kernel", "achieved": ach, "peak": peak * ng, "peak_source": peak_src +
                          " x %d GPUs" % ng, "unit": "GB/s", "frac": ach / (peak * ng), "alg_bytes_per_px_iter": 80, "traffic": None},
                clocks=clocks)
    print(json.dumps(line))
    g.close()


if __name__ == "__main__":
    main()
