"""The C++ host layer (sparsebase_b200/host/include): the reference's class names and call
signatures over the C ABI.  tests/cpp/host_api_test.cc mirrors the reference's own unit
tests for this path with the reference's golden vectors; this module builds and runs it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_api_test.cc")
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "host_api_test")
INC = os.path.join(ROOT, "sparsebase_b200", "host", "include")


def build_host_test(force=False):
    from sparsebase_b200 import build as b
    b.build()
    deps = [SRC, os.path.join(ROOT, "include", "sb200.h")]
    for d, _, files in os.walk(INC):
        deps += [os.path.join(d, f) for f in files]
    if (not force and os.path.exists(BIN)
            and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps)):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", f"-I{INC}", SRC, "-o", BIN,
           f"-L{os.path.join(ROOT, 'sparsebase_b200')}", "-lsb200",
           "-Wl,-rpath,$ORIGIN/../../../sparsebase_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return BIN


def test_host_layer_compiles_and_device_free_checks_pass():
    exe = build_host_test()
    r = subprocess.run([exe, "--no-device"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "0 failure(s)" in r.stdout


def test_reference_include_paths_exist():
    """User code keeps its #include lines: every reference header of the path has a
    same-named header here."""
    for rel in ["bases/reorder_base.h", "format/csr.h", "format/coo.h", "format/csc.h",
                "format/cuda_csr_cuda.cuh", "format/cuda_array_cuda.cuh",
                "context/cuda_context_cuda.cuh", "converter/converter_order_two.h",
                "reorder/degree_reorder.h", "reorder/rcm_reorder.h",
                "permute/permute_order_two.h", "permute/permute_order_one.h",
                "feature/degree_distribution.h", "feature/degrees.h",
                "utils/function_matcher_mixin.h", "utils/exception.h"]:
        assert os.path.exists(os.path.join(INC, "sparsebase", rel)), rel


@pytest.mark.gpu
def test_host_layer_on_device():
    exe = build_host_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-2000:]
    assert "0 failure(s)" in r.stdout
