"""One-off stress of the plan selection: random shapes well beyond the committed randomized test (d up to 72, K up to 140,
full and diagonal covariances; set MLF_FUZZ_WIDE=1 for d up to 128 and K up to 256), CUDA path vs the oracle with the parity tolerances of tests/test_emgmm_gpu.py.
Usage (GPU box): python scripts/fuzz_parity.py [n_cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mlfortran_b200 import mlf_algo_emgmm, synth
from oracle.emgmm_oracle import get

LL_RTOL, PAR_RTOL = 1e-10, 1e-9
rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
orc = get()
bad = skipped = 0
plans = {}
t0 = time.time()
for i in range(n_cases):
    wide = os.environ.get("MLF_FUZZ_WIDE") == "1"
    d = int(rng.choice([rng.integers(1, 9), rng.integers(9, 17), rng.integers(17, 33), rng.integers(33, 129 if wide else 73)]))
    K = int(rng.choice([rng.integers(1, 9), rng.integers(9, 65), rng.integers(65, 257 if wide else 141)]))
    diag = bool(rng.random() < 0.35)
    npc = int(rng.integers(max(2 * d + 2, 8), max(2 * d + 3, 60)))
    while npc * K * K * d * d > 1.5e9 and npc > 2 * d + 2:
        npc = max(2 * d + 2, npc // 2)
    if npc * K * K * d * d > 3e9:
        K = max(1, int((3e9 / (npc * d * d)) ** 0.5))
    sc = float(rng.choice([1.5, 3.0, 6.0]))
    mp = float(rng.choice([0.0, 0.3, 1.0]))
    iters = int(rng.integers(1, 4))
    seed = 5000 + i
    data = synth.make_mixture(d, K, npc, sc, 1.0, seed)
    X = data["X"]
    Mu0, Cov0, lam0 = synth.injected_init(data["centres"], seed)
    orc.set_cov_diag(diag)
    try:
        oMu, oCov, olam, otr = orc.em_iterations(X, Mu0, Cov0, lam0, mp, iters)
    finally:
        orc.set_cov_diag(False)
    if not (np.isfinite(oCov).all() and np.isfinite(otr).all()):
        skipped += 1
        continue
    em = mlf_algo_emgmm()
    if em.init(X, K, mp, cov_type="diag" if diag else "full") != 0:
        bad += 1
        print("INIT REFUSED", dict(d=d, K=K, diag=diag), flush=True)
        continue
    em.inject(Mu0, Cov0, lam0)
    plan = em.kernel_plan().split(":")[0]
    plans[plan] = plans.get(plan, 0) + 1
    info, _ = em.step(iters)
    tr = em.sumLL_trace()
    ok = info == 0 and tr.shape == otr.shape and np.abs((tr - otr) / otr).max() <= LL_RTOL and rel(em.Mu, oMu) <= PAR_RTOL \
        and rel(em.Cov, oCov) <= PAR_RTOL and np.abs(em.lambda_ - olam).max() <= PAR_RTOL
    if not ok:
        bad += 1
        print("FAIL", dict(d=d, K=K, npc=npc, sc=sc, mp=mp, iters=iters, diag=diag, seed=seed), plan, "info", info, em.last_error(),
              "ll", float(np.abs((tr - otr) / otr).max()) if tr.shape == otr.shape else None, "mu", rel(em.Mu, oMu), "cov", rel(em.Cov, oCov), flush=True)
    em.finalize()
print(f"fuzz: {n_cases} cases, {bad} failures, {skipped} skipped (oracle NaN), {time.time() - t0:.0f} s; plans: {plans}")
sys.exit(1 if bad else 0)
