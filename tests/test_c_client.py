"""A plain C program against the C-ABI (tests/c_client/emgmm_client.c, modelled on the reference's tests/test_kmeans_c.c):
compiled and linked on CPU (no compute), compiled and RUN on a GPU box."""
import os
import subprocess

import pytest

from mlfortran_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    _lib.load()
    exe = str(tmp_path / "emgmm_client")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_client", "emgmm_client.c"),
                           "-o", exe, "-L", libdir, "-lmlfortran_b200", f"-Wl,-rpath,{libdir}", "-lm"])
    return exe


def test_c_client_compiles_and_links(tmp_path):
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_c_client_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "emgmm_c_client.h5")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
