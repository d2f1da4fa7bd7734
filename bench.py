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
           every warp's iteration loop, faldoi_solver_last_iter_ms) vs the measured HBM copy bandwidth.
           `traffic` = DRAM bytes of ONE launch (ncu, 16-pair capture scaled to this batch).
  cpu_baseline: the UNMODIFIED reference tvl2OF (oracle/_ref, OpenMP, all host cores) on ONE pair of
           the same workload (about 5-10 s); falls back to the C port (oracle/) if the reference
           build did not travel.
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
    ap.add_argument("--mode", default="pairs", choices=["pairs", "stripes"],
                    help="pairs: independent pairs per GPU (default, the driver's line). stripes: ONE frame pair cut into "
                         "row stripes over --gpus GPUs with halo exchange over NVLink (3840x2160 by default); one host "
                         "process drives all GPUs (under torchrun rank 0 does, the other ranks exit)")
    return ap.parse_args()


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


# ----------------------------------------------------------------------------- CPU arms
def cpu_solve_one(I0, I1, u0, warps, method):
    """Times the reference's own CPU implementation on one pair.  Returns (seconds, kind, cores, u_out)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as po
    cores = os.cpu_count() or 1
    if po.have_ref():
        t = time.perf_counter()
        u, _ = po.r_global_solve(method, I0, I1, None, None, u0, None, warps=warps)
        return time.perf_counter() - t, "reference", cores, u
    t = time.perf_counter()
    u, _, _, _ = po.o_global_solve(method, I0, I1, None, None, u0, None, warps=warps)
    return time.perf_counter() - t, "port", cores, u


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
    if method in (2, 3, 6, 7):  # NLTV arithmetic: the reference's (bit-identical flows) unless FALDOI_NLTV_FAST=1
        cfg["nltv_arithmetic"] = "fast (approximate divisions, paired slots)" if os.environ.get("FALDOI_NLTV_FAST", "0") not in ("", "0") else "exact"
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
        import torch
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        d = make_pairs(1, w, h, 1000, dev)
        I0, I1, u0 = [d[k][0].cpu().numpy() for k in ("I0", "I1", "u0")]
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import pyoracle as po
        _, _, iters, _ = po.o_global_solve(method, I0, I1, None, None, u0, warps=a.warps) if method == 0 else (0, 0, [400] * a.warps, 0)
        times, kind, cores = [], None, None
        for k in range(a.warmup + a.steps):
            t, kind, cores, _ = cpu_solve_one(I0, I1, u0, a.warps, method)
            if k >= a.warmup:
                times.append(t)
        tot = sum(times)
        val = npix * sum(iters) * a.steps / tot / 1e6
        line = dict(base, impl="reference", value=val, ms_per_step=1e3 * tot / a.steps, gpu_launches=0,
                    cpu_baseline={"value": val, "unit": "Mpix*iter/s", "cores": cores, "kind": kind,
                                  "sample": "1 pair (seed 1000) of the workload per step, %d iterations" % sum(iters)},
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
    d = make_pairs(B, w, h, 1000 + rank * B, dev)
    need_lab = method in (2, 3, 6, 7)
    lab = None
    if need_lab:
        # Lab conversion is preprocessing (once, outside the timed region): the product's own host
        # library does it (libfaldoi_host.so, include/faldoi_host.h) -- not the oracle.
        import ctypes as C
        hostlib = C.CDLL(os.path.join(ROOT, "faldoi-ipol_b200", "libfaldoi_host.so"))
        rgb = np.ascontiguousarray(d["rgb"].cpu().numpy())
        lab_np = np.empty_like(rgb)
        for k in range(B):
            hostlib.faldoi_host_image_to_lab(rgb[k].ctypes.data_as(C.c_void_p), w, h, lab_np[k].ctypes.data_as(C.c_void_p))
        lab = torch.from_numpy(lab_np).to(dev)
    chi = torch.zeros(B, h, w, device=dev) if method == 8 else None
    host = {k: d[k].cpu().pin_memory() for k in ("I0", "I1", "Im1", "u0")}
    host_lab = lab.cpu().pin_memory() if lab is not None else None
    host_chi = chi.cpu().pin_memory() if chi is not None else None
    host_out = torch.empty(B, 2, h, w).pin_memory()
    host_chi_out = torch.empty(B, h, w).pin_memory() if method == 8 else None

    solver = fb.Solver(w, h, method, B, device=local_rank)
    params = fb.default_params(method, a.glb_iters, a.warps)
    stream = torch.cuda.ExternalStream(solver.stream, device=dev)

    def ptr(t, k):
        return t[k].data_ptr() if t is not None else 0

    def stage(src, lab_t, chi_t):
        for k in range(B):
            solver.upload_ptrs(k, ptr(src["I0"], k), ptr(src["I1"], k), ptr(src["u0"], k),
                               Im1=ptr(src["Im1"], k) if method == 8 else 0, lab=ptr(lab_t, k), chi=ptr(chi_t, k))

    def step_resident():
        stage(d, lab, chi)
        solver.run(params)

    def step_e2e():
        stage(host, host_lab, host_chi)
        solver.run(params)
        for k in range(B):
            solver.download_ptr(k, host_out[k].data_ptr(), host_chi_out[k].data_ptr() if host_chi_out is not None else 0)

    # warm-up (also yields the per-pair iteration counts: deterministic, identical every step)
    for _ in range(max(a.warmup, 1)):
        step_resident()
        solver.sync()
    iters_total = 0
    for k in range(B):
        _, _, log = solver.download(k)
        iters_total += sum(log.iters[:a.warps])
    units_step = npix * iters_total  # pixel-iterations per step on this rank

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed: resident
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    iter_ms, launches = 0.0, 0
    for _ in range(a.steps):
        step_resident()
        solver.sync()  # needed to read last_iter_ms; the stream stays back-to-back within a step
        iter_ms += solver.last_iter_ms
        launches += solver.last_launches
    e1.record(stream)
    barrier()
    t_res = e0.elapsed_time(e1) / 1e3
    # ---- timed: end to end with host buffers
    for _ in range(1):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop()

    tt = torch.tensor([t_res, t_e2e, iter_ms], device=dev, dtype=torch.float64)
    uu = torch.tensor([float(units_step), float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(uu, op=dist.ReduceOp.SUM)
    t_res, t_e2e, iter_ms_max = [float(x) for x in tt.cpu()]
    units_all, launches_all = [float(x) for x in uu.cpu()]

    if rank == 0:
        value = units_all * a.steps / t_res / 1e6
        e2e = units_all * a.steps / t_e2e / 1e6
        peak, peak_src = peak_hbm()
        alg = ALG_BYTES[method]
        # rank 0's own kernel: bytes it moved algorithmically / its time inside the iteration launches
        ach = alg * units_step * a.steps / (iter_ms / 1e3) / 1e9
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = json.load(open(prof)).get(str(method)) if os.path.exists(prof) else None
        if traffic is not None:
            traffic = traffic * B / 16.0  # the ncu capture had 16 pairs per launch; bytes scale with the pairs
        nplanes_in = 4 + (1 if method == 8 else 0) * 2 + (3 if need_lab else 0)
        line = dict(base, value=value, ms_per_step=1e3 * t_res / a.steps,
                    e2e={"value": e2e, "unit": "Mpix*iter/s", "h2d_bytes_per_step": B * nplanes_in * npix * 4,
                         "d2h_bytes_per_step": B * (2 + (1 if method == 8 else 0)) * npix * 4,
                         "pairs_per_s": world * B * a.steps / t_e2e},
                    pairs_per_s=world * B * a.steps / t_res, iters_per_pair=iters_total / B,
                    gpu_launches=int(launches_all),
                    roofline={"bound": "hbm", "kernel": {0: "tv_tile2_kernel", 1: "tv_tile2_kernel", 4: "tv_tile_kernel<CSAD>", 5: "tv_tile_kernel<CSAD>", 2: "nltv_tile_kernel", 3: "nltv_tile_kernel",
                                         6: "nltv_tile_kernel<CSAD>", 7: "nltv_tile_kernel<CSAD>", 8: "occ_xi_rows_kernel + occ_chi_rows_kernel"}[method],
                              "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                              "alg_bytes_per_px_iter": alg, "traffic": traffic},
                    clocks=clocks)
        if not a.no_cpu_baseline and world == 1 and method in (0, 1):
            k0 = 0
            I0, I1, u0 = [host[k][k0].numpy() for k in ("I0", "I1", "u0")]
            _, _, log = solver.download(k0)
            t, kind, cores, u_cpu = cpu_solve_one(I0, I1, u0, a.warps, method)
            its = sum(log.iters[:a.warps])
            line["cpu_baseline"] = {"value": npix * its / t / 1e6, "unit": "Mpix*iter/s", "cores": cores, "kind": kind,
                                    "sample": "pair 0 of the workload, %d iterations, %.2f s; flow bit-identical to GPU: %s"
                                              % (its, t, bool(np.array_equal(u_cpu, host_out[k0].numpy())))}
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


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
