"""Generates tests/golden/*.npz with the CPU oracle (LAPACK path).  The reference holds no golden vectors for the
EM path and cannot be run here, so these fixtures pin the *oracle* (regression guard) and give the GPU tests a
fixed target that does not depend on the oracle being rebuilt.  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mlfortran_b200 import synth  # noqa: E402
from oracle.emgmm_oracle import get  # noqa: E402

CASES = {
    # name: (d, K, n_per_comp, sigma_c, sigma_x, seed, min_proba, iters)
    "cfg1_test_emgmm": (2, 3, 1000, 5.0, 1.0, 20240601, 1.0, 49),   # tests/test_emgmm.f90 shape, `2 3 1000 50 5.0 1.0`
    "cfg1_free_lambda": (2, 3, 1000, 5.0, 1.0, 20240601, 0.0, 20),
    "d16_k8": (16, 8, 400, 8.0, 1.0, 3, 1.0, 8),
    "d5_k7_ragged": (5, 7, 143, 4.0, 1.0, 17, 0.0, 12),
}


def main():
    orc = get()
    assert orc.has_lapack
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (d, K, npc, sc, sx, seed, mp, iters) in CASES.items():
        data = synth.make_mixture(d, K, npc, sc, sx, seed)
        Mu0, Cov0, lam0 = synth.injected_init(data["centres"], seed)
        Mu, Cov, lam, trace, P = orc.em_iterations(data["X"], Mu0, Cov0, lam0, mp, iters, return_proba=True)
        np.savez_compressed(os.path.join(here, name + ".npz"), X=data["X"], Mu0=Mu0, Cov0=Cov0, lam0=lam0, min_proba=mp, iters=iters,
                            Mu=Mu, Cov=Cov, lam=lam, sumLL=trace, labels=orc.get_class(P), P_last_colsum=P.sum(1))
        print(name, data["X"].shape, "sumLL[-1] =", trace[-1])


if __name__ == "__main__":
    main()
