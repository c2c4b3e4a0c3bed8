"""The drop-in proof: the UNMODIFIED reference (built here with USE_CUDA from /root/reference,
binary shipped under oracle/_ref/) with sparsebase_b200/host/plugin/sb200_sparsebase_plugin.h
registered into its converter and operators.  oracle/plugin_demo.cc runs every operator of the
path through the reference's own dispatch twice -- CPUContext (reference code) and CUDAContext
(libsb200.so) -- and memcmp's the results."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "plugin_demo")


@pytest.mark.gpu
def test_reference_dispatch_reaches_sb200_and_results_match():
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/plugin_demo not built (needs the reference tree: make -C oracle ref)")
    r = subprocess.run([DEMO], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "ALL EQUAL" in r.stdout


@pytest.mark.gpu
def test_plugin_end_to_end_host_csr_in_host_csr_out():
    """oracle/plugin_bench.cc: RCMReorder + Permute2D through the reference's own API with a host
    format::CSR in and host arrays out (the plugin's staged CSR <-> CUDACSR transfers), checked
    against the reference's CPU functions in the same process."""
    import json
    exe = os.path.join(ROOT, "oracle", "_ref", "plugin_bench")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/plugin_bench not built (needs the reference tree: make -C oracle ref)")
    r = subprocess.run([exe, "700", "1", "--check"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["parity"] is True and line["nnz"] == 5 * 700 * 700 - 4 * 700
