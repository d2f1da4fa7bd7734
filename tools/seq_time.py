"""Times `global_faldoi -seq` on N copies of the full-size Sintel pair (oracle/_ref/data/clean_easy):
wall seconds per pair through the CLI, files in, files out."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "faldoi-ipol_b200", "bin", "global_faldoi")
D = os.path.join(ROOT, "oracle", "_ref", "data", "clean_easy")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
method = sys.argv[2] if len(sys.argv) > 2 else "0"
with tempfile.TemporaryDirectory() as t:
    ims = os.path.join(t, "ims.txt")
    open(ims, "w").write("".join(os.path.join(D, "frame_%04d.png" % k) + "\n" for k in (2, 3, 1, 4)))
    t1 = None
    for jobs in (1, 4, 4 + n):
        open(os.path.join(t, "jobs.txt"), "w").write("".join("%s %s %s\n" % (ims, os.path.join(D, "rg.flo"), os.path.join(t, "o%d.flo" % k)) for k in range(jobs)))
        t0 = time.perf_counter()
        r = subprocess.run([BIN, "-seq", os.path.join(t, "jobs.txt"), "-m", method, "-w", "5"], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        print("%d pairs: %.2f s wall" % (jobs, dt), flush=True)
        if jobs == 4:
            t1 = dt
    print("marginal: %.3f s per pair (%.1f pairs/s); 4 pairs incl. start-up %.2f s" % ((dt - t1) / n, n / (dt - t1), t1))
    same = open(os.path.join(t, "o0.flo"), "rb").read() == open(os.path.join(D, "var_m0.flo"), "rb").read() if method == "0" else None
    print("o0.flo identical to the reference's var_m0.flo:", same)
