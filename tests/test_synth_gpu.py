"""GPU checks of the device workload generator `mlf_b200_synth_gmm` (csrc/mlf_synth.cu) through the C-ABI:
agreement with the CPU restatement (oracle/synth_oracle.py) to rounding, shard independence, the statistics of the recipe of
tests/test_emgmm.f90:105-124 at a size the restatement could not reach, and an EM run on generated data that recovers the
generating mixture (what the reference's test program prints: Mu against XC, Cov against C12 C12^T)."""
import numpy as np
import pytest

from mlfortran_b200 import _lib, mlf_algo_emgmm, synth
from oracle import synth_oracle as so

pytestmark = pytest.mark.gpu

TOL = 1e-12   # Box-Muller through CUDA's log/sincospi vs libm's log/sin/cos, then a d-term dot product: rounding-level


def _gen_host(d, K, n0, n, N, sc, sx, seed):
    X = np.empty((n, d))
    c = np.empty((K, d))
    C12f = np.empty((K, d, d))
    rc = _lib.load().mlf_b200_synth_gmm(X.ctypes.data, n0, n, N, d, K, sc, sx, seed, 0, 0, c.ctypes.data, C12f.ctypes.data)
    assert rc == 0
    return X, c, C12f.transpose(0, 2, 1)


@pytest.mark.parametrize("d,K,N", [(1, 1, 5), (2, 3, 1000), (5, 4, 777), (16, 8, 4099), (33, 2, 300), (128, 3, 70)])
def test_device_points_match_restatement(d, K, N):
    X, c, C12 = _gen_host(d, K, 0, N, N, 6.0, 1.0, 100 + d)
    oX, _, oc, oC = so.points(d, K, 0, N, N, 6.0, 1.0, 100 + d)
    assert np.abs(c - oc).max() <= 1e-13 * max(1.0, np.abs(oc).max())
    assert np.abs(C12 - oC).max() <= 1e-13
    assert np.abs(X - oX).max() <= TOL * max(1.0, np.abs(oX).max())


def test_shards_are_slices_of_the_whole_and_device_buffers_agree():
    import torch

    d, K, N = 16, 8, 100003
    X, _, _ = _gen_host(d, K, 0, N, N, 8.0, 1.0, 3)
    for n0, n in [(0, 1), (1, 31), (32, 33), (50000, 50003), (N - 7, 7)]:
        Xs, _, _ = _gen_host(d, K, n0, n, N, 8.0, 1.0, 3)
        assert np.array_equal(Xs, X[n0:n0 + n])          # bit-identical: a pure function of (seed, global index)
    Xd, c, C12 = synth.make_mixture_device(d, K, 60000, 40003, N, 8.0, 1.0, 3, "cuda:0")
    assert np.array_equal(Xd.cpu().numpy(), X[40003:])
    X2, _, _ = _gen_host(d, K, 0, 1000, N, 8.0, 1.0, 4)   # keyed
    assert np.abs(X2 - X[:1000]).max() > 1.0


def test_statistics_of_the_recipe_at_scale():
    import torch

    d, K, N = 8, 16, 4_000_000
    X, c, C12 = synth.make_mixture_device(d, K, N, 0, N, 8.0, 1.0, 17, "cuda:0")
    src = so.permutation(17, N, np.arange(N, dtype=np.uint64))
    comp = torch.as_tensor((src % np.uint64(K)).astype(np.int64), device=X.device)
    cnt = torch.bincount(comp, minlength=K).cpu().numpy()
    assert cnt.max() - cnt.min() <= 1
    for k in range(K):
        Xk = X[comp == k]
        m = Xk.mean(0)
        S = ((Xk - m).T @ (Xk - m) / Xk.shape[0]).cpu().numpy()
        T = C12[k] @ C12[k].T
        se = np.sqrt(np.diag(T) / Xk.shape[0])
        assert np.all(np.abs(m.cpu().numpy() - c[k]) < 5 * se)
        assert np.abs(S - T).max() < 5 * np.abs(T).max() * np.sqrt(2.0 / Xk.shape[0]) * 2
    # neighbours in memory come from unrelated components (the shuffle of tests/test_emgmm.f90:122-124)
    same = (comp[1:] == comp[:-1]).double().mean().item()
    assert abs(same - 1.0 / K) < 0.005


def test_em_recovers_the_generating_mixture():
    """testGMM end to end on the device generator: EM from a jittered start ends at the generating centres and covariances."""
    d, K, N = 4, 5, 200000
    X, c, C12 = synth.make_mixture_device(d, K, N, 0, N, 8.0, 1.0, 23, "cuda:0")
    Mu0, Cov0, lam0 = synth.injected_init(c, 23)
    em = mlf_algo_emgmm()
    assert em.init(X, K, 1.0) == 0
    em.inject(Mu0, Cov0, lam0)
    info, _ = em.step(25)
    assert info == 0, em.last_error()
    assert np.abs(em.Mu - c).max() < 0.05
    for k in range(K):
        T = C12[k] @ C12[k].T
        assert np.abs(em.Cov[k] - T).max() < 0.05 * np.abs(T).max() + 0.02
    em.finalize()
