"""A clean build of the CUDA library from the sources alone (no prebuilt .so involved): what a fresh clone does.
nvcc cross-compiles for sm_100a without a GPU; the result must load and export the whole C ABI."""
import ctypes as C
import importlib
import os
import re
import subprocess

from conftest import ROOT


def test_clean_build_from_sources(tmp_path):
    bmod = importlib.import_module("faldoi-ipol_b200.build")
    out = str(tmp_path / "libfaldoi_gpu_clean.so")
    cmd = [bmod.NVCC] + [f for f in bmod.FLAGS if f not in ("-Xptxas", "-v")] + ["-o", out, os.path.join(bmod.CSRC, "faldoi_gpu.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lib = C.CDLL(out)
    txt = open(os.path.join(ROOT, "include", "faldoi_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    syms = sorted(set(re.findall(r"\b(faldoi_[a-zA-Z0-9_]+)\s*\(", txt)))
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), s
    # the kernels are sm_100a SASS with TMA tile loads
    sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    assert "sm_100a" in sass and sass.count("UTMALDG") >= 30
