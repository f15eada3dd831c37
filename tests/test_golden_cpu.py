"""The oracle against the committed golden fixtures (tests/golden/, made by make_golden.py with the LAPACK path),
with both eigen back-ends, and the independent NumPy restatement against the same fixtures."""
import glob
import os

import numpy as np
import pytest

from oracle import emgmm_numpy as onp

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
@pytest.mark.parametrize("lapack", [True, False])
def test_oracle_reproduces_golden(oracle, path, lapack):
    g = np.load(path)
    oracle.use_lapack(lapack)
    try:
        Mu, Cov, lam, tr, P = oracle.em_iterations(g["X"], g["Mu0"], g["Cov0"], g["lam0"], float(g["min_proba"]), int(g["iters"]), return_proba=True)
    finally:
        oracle.use_lapack(True)
    np.testing.assert_allclose(tr, g["sumLL"], rtol=1e-11)
    np.testing.assert_allclose(Mu, g["Mu"], rtol=0, atol=1e-9 * np.abs(g["Mu"]).max())
    np.testing.assert_allclose(Cov, g["Cov"], rtol=0, atol=1e-9 * np.abs(g["Cov"]).max())
    np.testing.assert_allclose(lam, g["lam"], rtol=0, atol=1e-10)
    assert (oracle.get_class(P) != g["labels"]).sum() == 0


def test_numpy_restatement_on_golden():
    g = np.load(GOLD[0])
    Mu, Cov, lam, tr, _ = onp.em_iterations(g["X"], g["Mu0"], g["Cov0"], g["lam0"], float(g["min_proba"]), 10)
    np.testing.assert_allclose(tr, g["sumLL"][:10], rtol=1e-11)
