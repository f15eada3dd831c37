"""Known-answer tests that pin the CPU oracle to the reference *formulas* (SURVEY.md section 8c, KAT-1..9).

The reference ships no golden vectors for the EM path (tests/test_emgmm.f90 only prints), so each
expected value below is derived analytically from the cited reference lines, and the C restatement
is cross-checked against the independent NumPy restatement.
"""
import numpy as np
import pytest

from mlfortran_b200 import synth
from oracle import emgmm_numpy as onp

LOG2PI = np.log(2.0 * np.arccos(-1.0))


def test_lapack_loaded(oracle):
    # the same dsytrd/dorgtr/dsteqr/dsyrk routines the reference calls (OpenBLAS bundled with SciPy)
    assert oracle.has_lapack


def test_kat1_single_component_uniform_weights(oracle):
    # KAT-1: K=1 => gamma=1; mu = sample mean; C = ms^2 * (1/N) sum (x-mu)(x-mu)^T with ms = N/(N-1)
    # (src/mlf_matrix.f90:119,137,145: ms multiplies B and C = B B^T, so it enters squared)
    rng = np.random.default_rng(1)
    N, d = 200, 3
    X = rng.standard_normal((N, d)) * [1.0, 2.0, 0.5] + [3.0, -1.0, 0.0]
    s, mu, Cm = oracle.max_gaussian(X, np.ones(N))
    assert s == pytest.approx(N, rel=0, abs=0)
    np.testing.assert_allclose(mu, X.mean(axis=0), rtol=1e-13)
    Z = X - X.mean(axis=0)
    expect = (N / (N - 1.0)) ** 2 * (Z.T @ Z) / N
    np.testing.assert_allclose(Cm, expect, rtol=1e-12)
    np.testing.assert_array_equal(Cm, Cm.T)


def test_kat2_identity_covariance_softmax(oracle):
    # KAT-2: C_k = I => lnDet = d/2 log 2pi under either sign => gamma = softmax_k(log lam_k - |x-mu_k|^2/2)
    rng = np.random.default_rng(2)
    N, d, K = 50, 4, 3
    X = rng.standard_normal((N, d))
    Mu = rng.standard_normal((K, d))
    lam = np.array([0.2, 0.5, 0.3])
    Cov = np.tile(np.eye(d), (K, 1, 1))
    info, P, ll = oracle.exp_step(X, Mu, Cov, lam)
    assert info == 0
    logit = np.log(lam)[:, None] - 0.5 * ((X[None] - Mu[:, None]) ** 2).sum(-1)
    e = np.exp(logit - logit.max(0))
    np.testing.assert_allclose(P, e / e.sum(0), rtol=1e-12)
    _, _, lnd = oracle.inverse_sym_matrix(np.eye(d))
    assert lnd == pytest.approx(0.5 * d * LOG2PI, rel=1e-15)


def test_kat3_lndet_sign_quirk(oracle):
    # KAT-3: C_k = s_k I: the eigen branch yields lnDet = d/2 log 2pi - (d/2) log s_k
    # (src/mlf_matrix.f90:285-290), so gamma_k is proportional to lam_k * s_k^{+d/2} * exp(-|x-mu_k|^2/(2 s_k)).
    d = 3
    s = np.array([0.25, 4.0])
    for sk in s:
        _, inv, lnd = oracle.inverse_sym_matrix(sk * np.eye(d))
        assert lnd == pytest.approx(0.5 * d * LOG2PI - 0.5 * d * np.log(sk), rel=1e-14)
        np.testing.assert_allclose(inv, np.eye(d) / sk, rtol=1e-14, atol=1e-300)
    X = np.array([[0.3, -0.2, 0.1], [1.5, 1.0, -2.0]])
    Mu = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, -1.0]])
    lam = np.array([0.6, 0.4])
    Cov = np.stack([sk * np.eye(d) for sk in s])
    _, P, _ = oracle.exp_step(X, Mu, Cov, lam)
    r2 = ((X[None] - Mu[:, None]) ** 2).sum(-1)
    unn = lam[:, None] * (s[:, None] ** (+d / 2.0)) * np.exp(-0.5 * r2 / s[:, None])
    np.testing.assert_allclose(P, unn / unn.sum(0), rtol=1e-12)
    wrong = lam[:, None] * (s[:, None] ** (-d / 2.0)) * np.exp(-0.5 * r2 / s[:, None])
    assert np.abs(P - wrong / wrong.sum(0)).max() > 1e-2  # the textbook sign would differ visibly


def test_kat4_singular_covariance_penalty(oracle):
    # KAT-4: zero eigenvalue -> replaced by 1/valM with valM = real(10e-9)*max eig (src/mlf_matrix.f90:283-289)
    Cm = np.diag([2.0, 0.0])
    info, inv, lnd = oracle.inverse_sym_matrix(Cm)
    assert info == 0
    valM = float(np.float32(10e-9)) * 2.0
    np.testing.assert_allclose(np.diag(inv), [0.5, 1.0 / valM], rtol=1e-14)
    assert lnd == pytest.approx(0.5 * (2 * LOG2PI + np.log(0.5) + np.log(1.0 / valM)), rel=1e-14)
    # duplicated coordinate: rank-1 covariance in 2-D
    Cm = np.array([[1.0, 1.0], [1.0, 1.0]])
    _, inv, _ = oracle.inverse_sym_matrix(Cm)
    valM = float(np.float32(10e-9)) * 2.0
    v1 = np.array([1.0, 1.0]) / np.sqrt(2)
    v0 = np.array([1.0, -1.0]) / np.sqrt(2)
    np.testing.assert_allclose(inv, 0.5 * np.outer(v1, v1) + (1 / valM) * np.outer(v0, v0), rtol=1e-9)


def test_kat5_threshold_drops_from_covariance_only(oracle):
    # KAT-5: gamma=(1,1,1,1e-5): W_4 ~ 3.3e-6 < 1e-3/4 => point 4 moves mu but not C (src/mlf_matrix.f90:132-138)
    X = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [100.0, 100.0]])
    g = np.array([1.0, 1.0, 1.0, 1e-5])
    s, mu, Cm = oracle.max_gaussian(X, g)
    W = g / g.sum()
    np.testing.assert_allclose(mu, W @ X, rtol=1e-14)
    ms = 1.0 / (1.0 - (W * W).sum())
    Z = X[:3] - mu
    expect = ms * ms * (Z.T * W[:3]) @ Z
    np.testing.assert_allclose(Cm, expect, rtol=1e-12)
    Zall = X - mu
    with_all = ms * ms * (Zall.T * W) @ Zall
    assert np.abs(Cm - with_all).max() > 1e-3
    # exactly at the threshold the point is kept ("W(i) < minW0 -> CYCLE")
    minW = 1e-3 / float(np.float32(4))
    g2 = np.array([1.0, 1.0, 1.0, 0.0])
    g2[3] = minW * 3.0 / (1.0 - minW)
    W2 = g2 / g2.sum()
    _, mu2, C2 = oracle.max_gaussian(X, g2)
    keep = ~(W2 < minW)
    ms2 = 1.0 / (1.0 - (W2 * W2).sum())
    Z2 = X[keep] - mu2
    np.testing.assert_allclose(C2, ms2 * ms2 * (Z2.T * W2[keep]) @ Z2, rtol=1e-12)


def test_kat6_minproba_floor(oracle):
    # KAT-6: lambda = max(lambda, minProba*max lambda)/sum (src/unsupervised/mlf_emgmm.f90:184-187)
    rng = np.random.default_rng(6)
    X = rng.standard_normal((60, 2))
    P = rng.random((3, 60)) * np.array([[1.0], [0.1], [0.01]])
    P /= P.sum(0)
    s = P.sum(1)
    _, _, lam0 = oracle.max_step(X, P, 0.0)
    np.testing.assert_allclose(lam0, s / s.sum(), rtol=1e-14)
    _, _, lam1 = oracle.max_step(X, P, 1.0)
    np.testing.assert_allclose(lam1, np.full(3, 1 / 3), rtol=1e-15)
    _, _, lamh = oracle.max_step(X, P, 0.5)
    f = np.maximum(s, 0.5 * s.max())
    np.testing.assert_allclose(lamh, f / f.sum(), rtol=1e-14)


def test_kat7_sumll_definition(oracle):
    # KAT-7: sumLL = sum_k [ N (log lam_k - lnDet_k) + sum_i (-1/2 maha_ik) ] -- unweighted over all pairs,
    # parameters at the start of the iteration (mlf_gaussian.f90:73-78, mlf_emgmm.f90:197-203)
    X = np.array([[0.0, 0.0], [1.0, 2.0], [-1.0, 0.5]])
    Mu = np.array([[0.0, 0.0], [1.0, 1.0]])
    Cov = np.stack([np.diag([1.0, 4.0]), np.diag([0.25, 1.0])])
    lam = np.array([0.7, 0.3])
    _, _, ll = oracle.exp_step(X, Mu, Cov, lam)
    exp = 0.0
    for k in range(2):
        dv = np.diag(Cov[k])
        lnDet = 0.5 * (2 * LOG2PI + np.log(1.0 / dv).sum())
        maha = (((X - Mu[k]) ** 2) / dv).sum(1)
        exp += 3 * (np.log(lam[k]) - lnDet) - 0.5 * maha.sum()
    assert ll == pytest.approx(exp, rel=1e-14)


def test_kat8_labels_first_max_one_based(oracle):
    P = np.array([[0.5, 0.2, 0.3], [0.5, 0.2, 0.4], [0.0, 0.6, 0.3]])  # (K=3, N=3)
    np.testing.assert_array_equal(oracle.get_class(P), [1, 3, 2])
    np.testing.assert_array_equal(onp.get_class(P), [1, 3, 2])
    Pn = np.full((2, 2), np.nan)
    Pn[1, 1] = 0.3
    np.testing.assert_array_equal(oracle.get_class(Pn), [1, 2])


@pytest.mark.parametrize("min_proba", [1.0, 0.0])
def test_c_vs_numpy_restatement(oracle, min_proba):
    # the two independent restatements agree on a test_emgmm-style workload over several iterations
    data = synth.make_mixture(d=3, K=4, n_per_comp=300, sigma_c=5.0, sigma_x=1.0, seed=11)
    Mu, Cov, lam = synth.injected_init(data["centres"], seed=11)
    a = oracle.em_iterations(data["X"], Mu, Cov, lam, min_proba, 6, return_proba=True)
    b = onp.em_iterations(data["X"], Mu, Cov, lam, min_proba, 6)
    np.testing.assert_allclose(a[3], b[3], rtol=1e-12)
    for u, v in zip(a[:3], b[:3]):
        np.testing.assert_allclose(u, v, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a[4], b[4], rtol=1e-8, atol=1e-14)
    np.testing.assert_array_equal(oracle.get_class(a[4]), onp.get_class(b[4]))


def test_lapack_vs_jacobi_paths(oracle):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((6, 6))
    Cm = A @ A.T + 0.1 * np.eye(6)
    oracle.use_lapack(True)
    _, inv_l, ln_l = oracle.inverse_sym_matrix(Cm)
    oracle.use_lapack(False)
    _, inv_j, ln_j = oracle.inverse_sym_matrix(Cm)
    oracle.use_lapack(True)
    np.testing.assert_allclose(inv_l, inv_j, rtol=1e-11, atol=1e-13)
    assert ln_l == pytest.approx(ln_j, rel=1e-13)
    np.testing.assert_allclose(inv_l @ Cm, np.eye(6), atol=1e-10)


def test_shipped_test_shape_converges(oracle):
    # config 1: the reference's own test program shape (ND=2 NC=3 NX=1000, minProba=1): means recovered
    data = synth.make_mixture(d=2, K=3, n_per_comp=1000, sigma_c=5.0, sigma_x=1.0, seed=20240601)
    Mu, Cov, lam = synth.injected_init(data["centres"], seed=20240601)
    Mu, Cov, lam, trace = oracle.em_iterations(data["X"], Mu, Cov, lam, 1.0, 49)
    assert np.abs(Mu - data["centres"]).max() < 0.2
    true_cov = np.einsum("kab,kcb->kac", data["C12"], data["C12"])
    assert np.abs(Cov - true_cov).max() < 0.25
    np.testing.assert_allclose(lam, 1 / 3, rtol=1e-15)


def test_diagonal_extension_c_vs_numpy(oracle):
    # EXTENSION (no reference semantics, SURVEY fact 0.8 / BASELINE configs[4]): Appendix A with every covariance restricted
    # to its diagonal.  The two independent restatements must agree, and the result must differ from the full-covariance run.
    data = synth.make_mixture(4, 3, 300, 4.0, 1.0, 33)
    X = data["X"]
    Mu0, Cov0, lam0 = synth.injected_init(data["centres"], 33)
    oracle.set_cov_diag(True)
    try:
        Mu, Cov, lam, tr = oracle.em_iterations(X, Mu0, Cov0, lam0, 0.0, 5)
    finally:
        oracle.set_cov_diag(False)
    nMu, nCov, nlam, ntr, _ = onp.em_iterations(X, Mu0, Cov0, lam0, 0.0, 5, cov_diag=True)
    np.testing.assert_allclose(tr, ntr, rtol=1e-11)
    np.testing.assert_allclose(Mu, nMu, rtol=0, atol=1e-10)
    np.testing.assert_allclose(Cov, nCov, rtol=0, atol=1e-10)
    off = Cov - np.stack([np.diag(np.diag(c)) for c in Cov])
    assert (off == 0).all()
    fMu, fCov, flam, ftr = oracle.em_iterations(X, Mu0, Cov0, lam0, 0.0, 5)
    assert np.abs(fCov - Cov).max() > 1e-3      # the full-covariance reference path is a different computation


@pytest.mark.parametrize("d,K,npc,sc,min_proba,seed", [(2, 3, 25, 3.0, 1.0, 11), (3, 3, 20, 2.0, 0.0, 5), (5, 2, 30, 1.0, 0.3, 7), (1, 4, 15, 4.0, 0.0, 9)])
def test_oracle_rounding_error_against_50_digit_arithmetic(oracle, d, K, npc, sc, min_proba, seed):
    """The fp64 oracle against the same formulas evaluated with 50 significant digits (oracle/emgmm_mp.py), iteration by
    iteration from the oracle's own fp64 state: sumLL to 1e-13 relative, parameters to 1e-12 -- the oracle's rounding error is
    three orders of magnitude inside the 1e-10 / 1e-9 budget the CUDA path is compared with."""
    from oracle import emgmm_mp

    data = synth.make_mixture(d, K, npc, sc, 1.0, seed)
    X = data["X"]
    Mu, Cov, lam = synth.injected_init(data["centres"], seed)
    for _ in range(3):
        mMu, mCov, mLam, mLL, _ = emgmm_mp.em_iteration(X, Mu, Cov, lam, min_proba)
        Mu, Cov, lam, tr = oracle.em_iterations(X, Mu, Cov, lam, min_proba, 1)
        eMu = np.array([[float(v) for v in r] for r in mMu])
        eCov = np.array([[[float(v) for v in r] for r in C] for C in mCov])
        eLam = np.array([float(v) for v in mLam])
        assert abs(float(mLL) - tr[0]) <= 1e-13 * abs(tr[0])
        assert np.abs(eMu - Mu).max() <= 1e-12 * max(1.0, np.abs(Mu).max())
        assert np.abs(eCov - Cov).max() <= 1e-12 * max(1.0, np.abs(Cov).max())
        assert np.abs(eLam - lam).max() <= 1e-13
