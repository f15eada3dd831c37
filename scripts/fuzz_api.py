"""One-off stress of the API paths around the iteration on random shapes: fresh object (k-means initialiser, then EM from its
result), getProba / getClass on new points, two shards in one process; everything against the oracle.
Usage (GPU box): python scripts/fuzz_api.py [n_cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mlfortran_b200 import mlf_algo_emgmm, synth
from oracle.emgmm_oracle import get

LL_RTOL, PAR_RTOL = 1e-10, 1e-9
rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
orc = get()
bad = 0
checked = {'em_after_kmeans': 0, 'getProba': 0, 'getClass': 0, 'two_shards': 0, 'streamed_refresh': 0, 'single_component': 0}
t0 = time.time()
for i in range(n_cases):
    d = int(rng.choice([rng.integers(1, 9), rng.integers(9, 17), rng.integers(17, 41)]))
    K = int(rng.choice([rng.integers(1, 9), rng.integers(9, 49)]))
    npc = int(rng.integers(max(2 * d + 2, 8), max(2 * d + 3, 80)))
    sc, mp, seed = float(rng.choice([3.0, 8.0])), float(rng.choice([0.0, 1.0])), 9000 + i
    data = synth.make_mixture(d, K, npc, sc, 1.0, seed)
    X = data["X"]
    tag = dict(d=d, K=K, N=X.shape[0], sc=sc, mp=mp, seed=seed)
    def fail(what, **kw):
        global bad
        bad += 1
        print("FAIL", what, tag, kw, flush=True)
        if os.environ.get("MLF_FUZZ_STOP") == "1":
            os._exit(1)
    # (a) fresh object: initialiser, then one EM iteration from whatever it produced
    em = mlf_algo_emgmm()
    if em.init(X, K, mp) != 0:
        fail("init"); continue
    em.set_seed(seed)
    if em.step(1)[0] != 0:
        fail("kmeans step", err=em.last_error()); em.finalize(); continue
    Mu1, Cov1, lam1 = em.Mu.copy(), em.Cov.copy(), em.lambda_.copy()
    if not (np.isfinite(Mu1).all() and np.isfinite(Cov1).all() and abs(lam1.sum() - 1) < 1e-12):
        fail("kmeans result not finite / lambda not normalised")
    oMu, oCov, olam, otr, oP = orc.em_iterations(X, Mu1, Cov1, lam1, mp, 1, return_proba=True)
    # a component owning (almost) a single point has ms^2 = 1/(1 - sum W^2)^2 ~ 1e17 in front of a cancelled outer product:
    # the reference's own result is rounding noise there (both sides are ~1e-7 from the exact value), nothing to compare
    Wn = oP / oP.sum(1, keepdims=True)
    wellc = (1.0 - (Wn * Wn).sum(1)) > 1e-4
    if np.isfinite(oCov).all() and np.isfinite(otr).all():
        checked['em_after_kmeans'] += 1
        if em.step(1)[0] != 0:
            fail("em step", err=em.last_error())
        elif not (abs(em.sumLL - otr[-1]) <= LL_RTOL * abs(otr[-1]) and rel(em.Mu, oMu) <= PAR_RTOL and rel(em.Cov[wellc], oCov[wellc]) <= PAR_RTOL
                  and np.abs(em.lambda_ - olam).max() <= PAR_RTOL):
            fail("em after kmeans", ll=abs(em.sumLL - otr[-1]) / abs(otr[-1]), mu=rel(em.Mu, oMu), cov=rel(em.Cov, oCov))
        # (b) inference on new points
        Q = np.ascontiguousarray(X[rng.permutation(X.shape[0])[: min(257, X.shape[0])]] + 0.1 * rng.standard_normal((min(257, X.shape[0]), d)))
        info, P = em.getProba(Q)
        _, Po, _ = orc.exp_step(Q, em.Mu, em.Cov, em.lambda_)
        checked['getProba'] += int(np.isfinite(Po).all())
        if info != 0 or not np.isfinite(Po).all() or np.abs(P - Po).max() > 1e-9:
            if np.isfinite(Po).all():
                fail("getProba", info=info, err=float(np.abs(P - Po).max()))
        info, cl = em.getClass(Q)
        if np.isfinite(Po).all():
            checked['getClass'] += 1
            ocl = orc.get_class(Po)
            top2 = np.sort(Po, axis=0)[-2:]
            tie = (top2[1] - top2[0]) <= 1e-9 * top2[1] if Po.shape[0] > 1 else np.zeros(Po.shape[1], bool)
            if info != 0 or not ((cl == ocl) | tie).all():
                fail("getClass", info=info, n=int(((cl != ocl) & ~tie).sum()))
    em.finalize()
    # (c) two shards in one process, injected parameters
    Mu0, Cov0, lam0 = synth.injected_init(data["centres"], seed)
    oMu, oCov, olam, otr = orc.em_iterations(X, Mu0, Cov0, lam0, mp, 2)
    if np.isfinite(oCov).all() and np.isfinite(otr).all() and X.shape[0] >= 4:
        em = mlf_algo_emgmm()
        if em.init(X, K, mp, devices=[0, 0]) != 0:
            fail("multi init"); continue
        em.inject(Mu0, Cov0, lam0)
        checked['two_shards'] += 1
        if em.step(2)[0] != 0:
            fail("multi step", err=em.last_error())
        else:
            tr = em.sumLL_trace()
            if not (np.abs((tr - otr) / otr).max() <= LL_RTOL and rel(em.Mu, oMu) <= PAR_RTOL and rel(em.Cov, oCov) <= PAR_RTOL):
                fail("two shards", ll=float(np.abs((tr - otr) / otr).max()), mu=rel(em.Mu, oMu), cov=rel(em.Cov, oCov))
        em.finalize()
# (d) refresh_x + step: the upload is chunked underneath the sweep (ragged sizes around the chunk boundaries), full and diagonal
for i in range(max(n_cases // 10, 4)):
    d, K = int(rng.integers(1, 17)), int(rng.integers(2, 33))
    diag = bool(rng.random() < 0.3)
    if diag and rng.random() < 0.5:
        d, K = int(rng.integers(1, 9)), int(rng.integers(128, 200))
    n = int(rng.integers(70000, 260000))
    data = synth.make_mixture(d, K, n // K + 1, 6.0, 1.0, 9500 + i)
    X1 = np.ascontiguousarray(data["X"][:n])
    X2 = np.ascontiguousarray(X1 + 0.03 * rng.standard_normal(X1.shape) + 0.1)
    Mu0, Cov0, lam0 = synth.injected_init(data["centres"], 9500 + i)
    tag = dict(d=d, K=K, N=n, diag=diag)
    em = mlf_algo_emgmm()
    assert em.init(X1, K, 1.0, cov_type="diag" if diag else "full") == 0
    em.inject(Mu0, Cov0, lam0)
    assert em.step(1)[0] == 0
    Mu1, Cov1, lam1 = em.Mu.copy(), em.Cov.copy(), em.lambda_.copy()
    assert em.refresh_x(X2) == 0
    info, _ = em.step(2)
    orc.set_cov_diag(diag)
    try:
        oMu, oCov, olam, otr = orc.em_iterations(X2, Mu1, Cov1, lam1, 1.0, 2)
    finally:
        orc.set_cov_diag(False)
    checked['streamed_refresh'] += 1
    tr = em.sumLL_trace()
    if info != 0 or not (np.abs((tr - otr) / otr).max() <= LL_RTOL and rel(em.Mu, oMu) <= PAR_RTOL and rel(em.Cov, oCov) <= PAR_RTOL):
        bad += 1
        print("FAIL streamed refresh", tag, em.kernel_plan()[:40], float(np.abs((tr - otr) / otr).max()), rel(em.Mu, oMu), rel(em.Cov, oCov), flush=True)
    em.finalize()
# (e) single-component entry points over all supported dimensions
from mlfortran_b200 import mlf_evalGaussian, mlf_maxGaussian
for i in range(max(n_cases // 4, 10)):
    d = int(rng.choice([rng.integers(1, 17), rng.integers(17, 129)]))
    n = int(rng.integers(2 * d + 3, 3000))
    A_ = rng.standard_normal((d, d))
    Cm = A_ @ A_.T / d + 0.3 * np.eye(d)
    mu = rng.standard_normal(d) * 3
    Xs = np.ascontiguousarray(rng.standard_normal((n, d)) @ np.linalg.cholesky(Cm).T + mu)
    lam = float(rng.random() + 0.05)
    info, P, ll = mlf_evalGaussian(Xs, mu, Cm, lam)
    oinfo, oP, oll = orc.eval_gaussian(Xs, mu, Cm, lam)
    checked['single_component'] += 1
    if info != 0 or oinfo != 0 or not np.allclose(P, oP, rtol=1e-10, atol=1e-300) or abs(ll - oll) > LL_RTOL * abs(oll):
        bad += 1
        print("FAIL evalGaussian", dict(d=d, n=n), info, float(np.abs(P / oP - 1).max()), abs(ll - oll) / abs(oll), flush=True)
    w = rng.random(n) ** int(rng.integers(1, 10))
    s_, m_, C_ = mlf_maxGaussian(Xs, w)
    os_, om, oC = orc.max_gaussian(Xs, w)
    if abs(s_ - os_) > 1e-12 * os_ or np.abs(m_ - om).max() > PAR_RTOL * np.abs(om).max() or np.abs(C_ - oC).max() > PAR_RTOL * np.abs(oC).max():
        bad += 1
        print("FAIL maxGaussian", dict(d=d, n=n), abs(s_ - os_) / os_, np.abs(m_ - om).max() / np.abs(om).max(), np.abs(C_ - oC).max() / np.abs(oC).max(), flush=True)
print(f"fuzz_api: {n_cases} cases, {bad} failures, checks done {checked}, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
