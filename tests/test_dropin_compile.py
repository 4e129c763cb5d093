"""The drop-in claim, checked: the reference's own call sites -- test/simple_test.cpp, test/gpu_vs_cpu_test.cpp and
bench/benchmark.cpp -- compile UNMODIFIED against the C++ mirror (host/ivf_flat_index.h through the forwarding
headers host/dropin/engine/*.h) and link against libvdb_b200.so.  The sources are read from /root/reference at test
time (never copied into the repository); the binaries land in host/dropin/_bin/ (git-ignored, they travel to the
GPU box), where the gpu-marked test runs them."""
import importlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-acceleratedvectordatabaseengine_b200")
DROPIN = os.path.join(PKG, "host", "dropin")
BIN = os.path.join(DROPIN, "_bin")
REF = "/root/reference"
PROGRAMS = {"simple_test": "test/simple_test.cpp", "gpu_vs_cpu_test": "test/gpu_vs_cpu_test.cpp",
            "benchmark": "bench/benchmark.cpp"}
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")


@pytest.mark.parametrize("name", sorted(PROGRAMS))
def test_reference_call_sites_compile_unmodified(name):
    src = os.path.join(REF, PROGRAMS[name])
    if not os.path.exists(src):
        pytest.skip("reference checkout absent (GPU box): the binaries were built in the build container")
    pkg.build()
    os.makedirs(BIN, exist_ok=True)
    # the source is piped in, so its `#include "../engine/ivf_flat_index.h"` resolves against the working directory
    # (host/dropin/test/ -> host/dropin/engine/, the forwarding headers), not against /root/reference/engine/
    cwd = os.path.join(DROPIN, os.path.dirname(PROGRAMS[name]))
    cmd = ["g++", "-std=c++17", "-O2", "-x", "c++", "-I", "/usr/local/cuda/include", "-I", os.path.join(ROOT, "include"),
           "-o", os.path.join(BIN, name), "-", "-L", PKG, "-lvdb_b200_storage", "-lvdb_b200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lpthread", "-Wl,-rpath,$ORIGIN/../../.."]
    with open(src, "rb") as f:
        r = subprocess.run(cmd, stdin=f, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.exists(os.path.join(BIN, name))


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [("simple_test", []), ("gpu_vs_cpu_test", ["20000", "32", "64", "32"]),
                                       ("benchmark", ["50000", "64", "64", "8"])])
def test_reference_programs_run_on_the_b200_path(name, args):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("host/dropin/_bin not built (needs the reference checkout at build time)")
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    out = r.stdout + r.stderr
    assert "failed" not in out.lower() or "0 failed" in out.lower(), out[-2000:]
