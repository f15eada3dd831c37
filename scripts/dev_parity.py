"""Development helper: CUDA path vs oracle on a few shapes; prints the deviations."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mlfortran_b200 import mlf_algo_emgmm, synth
from oracle.emgmm_oracle import get

orc = get()

def run(d, K, npc, sc, sx, seed, min_proba, iters):
    data = synth.make_mixture(d, K, npc, sc, sx, seed)
    X = data["X"]
    Mu0, Cov0, lam0 = synth.injected_init(data["centres"], seed)
    t0 = time.time()
    oMu, oCov, olam, otr, oP = orc.em_iterations(X, Mu0, Cov0, lam0, min_proba, iters, return_proba=True)
    t1 = time.time()
    em = mlf_algo_emgmm()
    assert em.init(X, K, min_proba) == 0
    em.inject(Mu0, Cov0, lam0)
    info, dt = em.step(iters)
    assert info == 0, (info, em.last_error())
    tr = em.sumLL_trace()
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    P = em.ProbaC
    print(f"d={d} K={K} N={X.shape[0]} minProba={min_proba} iters={iters}: "
          f"sumLL rel {np.abs((tr - otr) / otr).max():.2e}  Mu {rel(em.Mu, oMu):.2e}  Cov {rel(em.Cov, oCov):.2e} "
          f"lambda {rel(em.lambda_, olam):.2e}  P(abs) {np.abs(P - oP).max():.2e}  oracle {t1-t0:.2f}s gpu {dt:.3f}s")
    lab = em.labels()
    info, Pq = em.getProba(X[:1000])
    _, Pn, _ = orc.exp_step(X[:1000], em.Mu, em.Cov, em.lambda_)
    print("   labels mismatches vs oracle argmax(final params):", int((lab[:1000] != orc.get_class(Pn)).sum()),
          " getProba abs diff", np.abs(Pq - Pn).max())
    em.finalize()

run(2, 3, 1000, 5.0, 1.0, 20240601, 1.0, 10)
run(2, 3, 1000, 5.0, 1.0, 20240601, 0.0, 10)
run(3, 4, 333, 5.0, 1.0, 11, 1.0, 5)
run(16, 8, 2000, 8.0, 1.0, 3, 1.0, 5)
run(16, 64, 500, 8.0, 1.0, 3, 1.0, 3)
run(7, 5, 1001, 4.0, 1.0, 5, 0.0, 6)
run(32, 6, 800, 4.0, 1.0, 6, 1.0, 3)
run(2, 8, 20000, 6.0, 1.0, 2, 0.0, 5)
