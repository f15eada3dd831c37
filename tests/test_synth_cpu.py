"""CPU checks of the workload generator: the oracle restatement (oracle/synth_oracle.py) against published
known answers and against the recipe of tests/test_emgmm.f90:105-124, and the host part of `mlf_b200_synth_gmm` (centres and
C12 need no GPU) against that restatement.  The point kernel itself is covered by tests/test_synth_gpu.py."""
import numpy as np
import pytest

from mlfortran_b200 import _lib, synth
from oracle import synth_oracle as so


def test_philox4x32_10_known_answers():
    """Random123's kat_vectors for philox4x32-10 (Salmon et al., SC'11): zero, all-ones and pi-digit inputs."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in so.philox4x32_10(*ctr, *key)) == want


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 16, 17, 1000, 4096, 4097, 100003])
def test_permutation_is_a_bijection(n):
    p = so.permutation(11, n, np.arange(n))
    assert np.array_equal(np.sort(p), np.arange(n, dtype=np.uint64))
    if n >= 1000:       # a shuffle, not the identity, and keyed
        assert (p != np.arange(n)).mean() > 0.99
        assert (so.permutation(12, n, np.arange(n)) != p).mean() > 0.99


@pytest.mark.parametrize("n", [1, 2, 3, 8, 15, 16])
def test_rand_orthogonal_properties(n):
    """randOrthogonal (src/mlf_matrix.f90:153-180): orthogonal, determinant +1 for either parity of n, columns scaled by Sigma."""
    H = so.rand_orthogonal(5, 0, n)
    assert np.abs(H.T @ H - np.eye(n)).max() < 1e-13
    assert abs(np.linalg.det(H) - 1.0) < 1e-12
    sig = np.linspace(1.0, 2.0, n)
    Hs = so.rand_orthogonal(5, 0, n, sig)
    assert np.abs(Hs - H * sig[None, :]).max() < 1e-15
    assert np.abs(so.rand_orthogonal(5, 1, n) - H).max() > 1e-3 or n == 1


def test_spectrum_follows_the_test_program():
    """weight = sqrt(sigmaX (1/i) / Mean(1/i)) (tests/test_emgmm.f90:116-117): the covariance trace is d sigmaX."""
    w = so.spectrum(7, 2.5)
    assert np.isclose((w ** 2).sum(), 7 * 2.5, rtol=1e-14)
    assert np.allclose(w, synth.spectrum(7, 2.5), rtol=1e-15)


@pytest.mark.parametrize("d,K", [(1, 1), (2, 3), (5, 4), (16, 8), (33, 2)])
def test_library_host_part_matches_restatement(d, K):
    """centres and C12 of mlf_b200_synth_gmm (host C++) against the NumPy restatement: same stream, same arithmetic order."""
    c, C12 = synth.mixture_params(d, K, 5.0, 1.5, 42 + d)
    oc, oC = so.mixture_params(d, K, 5.0, 1.5, 42 + d)
    assert np.abs(c - oc).max() <= 1e-13 * max(1.0, np.abs(oc).max())
    assert np.abs(C12 - oC).max() <= 1e-13
    w = so.spectrum(d, 1.5)
    for k in range(K):     # C12 = O diag(w)
        O = C12[k] / w[None, :]
        assert np.abs(O.T @ O - np.eye(d)).max() < 1e-12


def test_restatement_distribution():
    """Points of the restatement follow the recipe: equal component counts (+-1), component means = centres, component
    covariances = C12 C12^T within sampling error."""
    d, K, N = 3, 4, 40002
    X, lab, c, C12 = so.points(d, K, 0, N, N, 6.0, 1.0, 9)
    cnt = np.bincount(lab - 1, minlength=K)
    assert cnt.max() - cnt.min() <= 1 and cnt.sum() == N
    for k in range(K):
        Xk = X[lab == k + 1]
        se = np.sqrt(np.diag(C12[k] @ C12[k].T) / len(Xk))
        assert np.all(np.abs(Xk.mean(0) - c[k]) < 5 * se)
        S = np.cov(Xk.T, bias=True)
        assert np.abs(S - C12[k] @ C12[k].T).max() < 0.08
    # shards are slices of the whole
    Xa, la, _, _ = so.points(d, K, 1000, 500, N, 6.0, 1.0, 9)
    assert np.array_equal(Xa, X[1000:1500]) and np.array_equal(la, lab[1000:1500])


def test_argument_errors():
    lib = _lib.load()
    buf = np.empty(8)
    assert lib.mlf_b200_synth_gmm(None, 0, 0, 10, 0, 2, 1.0, 1.0, 1, 0, 0, None, None) < 0      # nY < 1
    assert lib.mlf_b200_synth_gmm(None, 0, 0, 10, 2, 0, 1.0, 1.0, 1, 0, 0, None, None) < 0      # nC < 1
    assert lib.mlf_b200_synth_gmm(buf.ctypes.data, 8, 4, 10, 2, 2, 1.0, 1.0, 1, 0, 0, None, None) < 0   # shard past the job
    assert lib.mlf_b200_synth_gmm(None, 0, 4, 10, 2, 2, 1.0, 1.0, 1, 0, 0, None, None) < 0      # points asked, no buffer
    assert lib.mlf_b200_synth_gmm(None, 0, 0, 10, 2, 2, 1.0, 1.0, 1, 0, 0, None, None) == 0     # parameters only: fine


def test_points_fail_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    buf = np.empty(8)
    assert _lib.load().mlf_b200_synth_gmm(buf.ctypes.data, 0, 4, 4, 2, 2, 1.0, 1.0, 1, 0, 0, None, None) == -7   # no CPU fallback
