"""The reference's own C++ wrapper (include/mlf.hpp + src/cpp_intf/mlfcpp.cpp, compiled unmodified from /root/reference)
against libmlfortran_b200.so -- SURVEY 2.1: MlfStepObject / MlfObject::getRsc / MlfHdf5 "must keep working".

CPU: compile + link here (where /root/reference exists), and every mlf_* symbol the wrapper needs must be exported by the
library.  GPU: run the binary built here (it travels with the snapshot; the reference's sources do not)."""
import os
import subprocess

import pytest

from mlfortran_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp_client", "_bin", "ref_wrapper_client")
REF = os.environ.get("MLF_REFERENCE", "/root/reference")


def test_reference_cpp_wrapper_links_against_the_library():
    if not os.path.exists(os.path.join(REF, "include", "mlf.hpp")):
        pytest.skip("reference sources not present (GPU box): the binary built in the container is used as is")
    _lib.load()
    subprocess.check_call(["bash", os.path.join(HERE, "cpp_client", "build.sh")])
    need = {ln.split()[-1] for ln in subprocess.check_output(["nm", "-u", EXE], text=True).splitlines() if " mlf_" in ln}
    have = {ln.split()[-1] for ln in subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True).splitlines() if " mlf_" in ln}
    assert {"mlf_step", "mlf_getrsc", "mlf_getinfo", "mlf_getnumrsc", "mlf_pushState", "mlf_hdf5_createFile", "mlf_hdf5_openFile", "mlf_dealloc"} <= need
    assert need <= have, f"unresolved against libmlfortran_b200.so: {sorted(need - have)}"


@pytest.mark.gpu
def test_reference_cpp_wrapper_runs(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("tests/cpp_client/_bin/ref_wrapper_client was not built (needs /root/reference at build time)")
    r = subprocess.run([EXE, str(tmp_path / "ref_wrapper.h5")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("niter;cpuTime;realTime;sumLL\n10;0;")
    assert "all checks passed" in r.stdout
