"""The monomial table of the quadratic-form sweep (qf_feature in mlfortran_b200/csrc/emgmm_kernels.cuh) is shared by the kernel that
builds the coefficient pack (k_refresh) and the sweep that forms the monomials (k_em16): compiled here for the host with g++ and
checked for what both rely on -- every linear term and every unordered pair of the padded coordinates exactly once, padding
features pointing at the zero slot, the k-step count of the header comment (38 at d = 16)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include <cstdio>
#include "emgmm_kernels.cuh"
using namespace mlfb200;
int main() {
  for (int DM = 1; DM <= 4; ++DM) {
    const int D4 = 4 * DM, NF = qf_nsteps(DM);
    printf("DM %d nfeat %d nsteps %d stride %d\n", DM, qf_nfeat(DM), NF, qf_stride(DM));
    for (int f = 0; f < 4 * NF; ++f) { int i = -1, j = -1; qf_feature(D4, f, i, j); printf("%d %d %d\n", f, i, j); }
  }
  return 0;
}
"""


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"), reason="needs g++ and the CUDA headers")
def test_monomial_table_covers_every_term_once(tmp_path):
    src = tmp_path / "qf.cpp"
    src.write_text(SRC)
    exe = tmp_path / "qf"
    subprocess.check_call(["g++", "-std=c++17", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "mlfortran_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    pos = 0
    for DM in (1, 2, 3, 4):
        hdr = out[pos].split(); pos += 1
        assert int(hdr[1]) == DM
        nfeat, nsteps, stride = int(hdr[3]), int(hdr[5]), int(hdr[7])
        D4 = 4 * DM
        assert nfeat == D4 + D4 * (D4 + 1) // 2 and nsteps == (nfeat + 3) // 4 and stride == nsteps * 32 + 8
        seen = {}
        for f in range(4 * nsteps):
            ff, i, j = (int(v) for v in out[pos].split()); pos += 1
            assert ff == f
            if f < nfeat:
                assert 0 <= i < D4 and i <= j <= D4            # j == D4 is the constant-1 slot (linear term)
                assert (i, j) not in seen
                seen[(i, j)] = f
            else:
                assert i == D4 + 1 and j == D4 + 1               # padding: the zero slot, coefficient 0
        assert len(seen) == nfeat
        assert all((i, D4) in seen for i in range(D4))           # every linear term
        assert all((i, j) in seen for i in range(D4) for j in range(i, D4))   # every unordered pair
    assert (4 * 4 + 16 * 17 // 2 + 3) // 4 == 38                 # d = 16: 152 monomials = 38 DMMA k-steps
