"""Row-stripe decomposition: partition logic on the CPU (incl. a world_size-2 gloo
run of the host-side sharding used by bench.py), and on the GPU the striped solve against the
whole-frame solve -- it must be bit-identical, iteration counts included."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, synthetic_pair


@pytest.mark.parametrize("h,n", [(436, 1), (436, 2), (436, 3), (2160, 8), (17, 8), (64, 5)])
def test_stripe_rows_partition(fb, h, n):
    rows = [fb.stripe_rows(h, n, k) for k in range(n)]
    assert rows[0][0] == 0 and rows[-1][1] == h
    assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    sizes = [r1 - r0 for r0, r1 in rows]
    assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


GLOO_WORKER = r'''
import os, sys, importlib
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
fb = importlib.import_module("faldoi-ipol_b200")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# pair-level sharding as bench.py does it: rank r takes pairs [r*B, (r+1)*B) of the sequence
B, total = 3, 3 * world
mine = list(range(rank * B, (rank + 1) * B))
got = [None] * world
dist.all_gather_object(got, mine)
assert sorted(sum(got, [])) == list(range(total))
# stripe partition of one frame: every rank derives the same plan, rows tile the frame
h = 437
plan = [fb.stripe_rows(h, world, k) for k in range(world)]
r0, r1 = plan[rank]
t = torch.zeros(h, dtype=torch.int32)
t[r0:r1] = 1
dist.all_reduce(t)
assert int(t.min()) == 1 and int(t.max()) == 1
# max-over-ranks timing reduction used by bench.py
x = torch.tensor([float(rank + 1)])
dist.all_reduce(x, op=dist.ReduceOp.MAX)
assert float(x) == world
dist.destroy_process_group()
print("ok", rank)
'''


def test_sharding_logic_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n", [(160, 70, 1), (300, 131, 1), (160, 70, 2), (160, 70, 3), (96, 33, 4), (300, 131, 5)])
def test_striped_equals_whole_frame_one_gpu(fb, po, w, h, n):
    """All stripes on device 0: exercises halo rows, peer stores, the cross-stripe ordering and the global error
    gather without needing several GPUs.  n = 1 runs every speculative block as ONE dataflow grid (tiles of
    successive launches waiting for each other); stripes that share a GPU get one grid per launch."""
    I0, I1, _, u0, _ = synthetic_pair(w, h, seed=w * 7 + h + n)
    p = fb.default_params(0, warps=3)
    whole, _, its, errs = fb.global_solve(0, I0, I1, u0, params=p)
    g = fb.Stripes(w, h, [0] * n)
    g.upload(I0, I1, u0)
    g.run(p)
    u, log = g.download()
    assert list(log.iters[:3]) == its and list(log.err[:3]) == errs
    assert np.array_equal(u, whole)
    ou, _, oits, _ = po.o_global_solve(0, I0, I1, None, None, u0, warps=3)
    assert its == oits and np.array_equal(u, ou)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nstripes", [1, 3])
def test_striped_4k_equals_whole_frame(fb, nstripes):
    """The shape the stripes exist for: a 3840x2160 pair in 3 stripes (device 0, or one per GPU when there are
    several) against the whole-frame solve, one warp; and as a single stripe = dataflow grids of up to 32
    launches over 7680 tiles."""
    w, h = 3840, 2160
    I0, I1, _, u0, _ = synthetic_pair(w, h, seed=9)
    p = fb.default_params(0, warps=1)
    whole, _, its, errs = fb.global_solve(0, I0, I1, u0, params=p)
    ndev = fb.device_count()
    g = fb.Stripes(w, h, [k % ndev for k in range(nstripes)])
    g.upload(I0, I1, u0)
    g.run(p)
    u, log = g.download()
    assert list(log.iters[:1]) == its and list(log.err[:1]) == errs
    assert np.array_equal(u, whole)
    g.close()


@pytest.mark.gpu
def test_striped_across_gpus(fb):
    """With >= 2 GPUs: one stripe per GPU, NVLink peer stores."""
    n = min(fb.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    w, h = 512, 300
    I0, I1, _, u0, _ = synthetic_pair(w, h, seed=5)
    p = fb.default_params(0, warps=2)
    whole, _, its, _ = fb.global_solve(0, I0, I1, u0, params=p)
    g = fb.Stripes(w, h, list(range(n)))
    g.upload(I0, I1, u0)
    g.run(p)
    u, log = g.download()
    assert list(log.iters[:2]) == its
    assert np.array_equal(u, whole)
    g.close()


def test_reference_arm_under_torchrun_uses_all_cores(po):
    """bench.py --impl reference launched the way the driver launches it for N > 1: torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers; rank 0 must pin the reference's OpenMP team back to every host core and be the
    only rank that prints a line (round 1's N>1 reference arm ran single-threaded: ratios inflated 2.9x)."""
    import json
    if not po.have_ref():
        pytest.skip("oracle/_ref not built here")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29543", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--width", "256", "--height", "128"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["omp_num_threads"] == os.cpu_count() and d["cpu_baseline"]["cores"] == os.cpu_count()
    assert d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
