"""Deviation of one CUDA EM iteration from the reference's formulas in 50-digit arithmetic (oracle/emgmm_mp.py), printed per
shape.  Usage (GPU box): python scripts/dev_exact.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mlfortran_b200 import mlf_algo_emgmm, synth
from oracle import emgmm_mp
from oracle.emgmm_oracle import get

rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)
orc = get()
for d, K, npc, sc, mp_, seed in [(2, 3, 25, 3.0, 1.0, 11), (3, 3, 20, 2.0, 0.0, 5), (5, 2, 30, 1.0, 0.3, 7), (8, 4, 50, 4.0, 1.0, 13), (16, 3, 40, 6.0, 1.0, 17)]:
    data = synth.make_mixture(d, K, npc, sc, 1.0, seed)
    X = data["X"]
    Mu0, Cov0, lam0 = synth.injected_init(data["centres"], seed)
    mMu, mCov, mLam, mLL, _ = emgmm_mp.em_iteration(X, Mu0, Cov0, lam0, mp_)
    eMu = np.array([[float(v) for v in r] for r in mMu]); eCov = np.array([[[float(v) for v in r] for r in C] for C in mCov]); eLam = np.array([float(v) for v in mLam])
    em = mlf_algo_emgmm(); assert em.init(X, K, mp_) == 0; em.inject(Mu0, Cov0, lam0); assert em.step(1)[0] == 0
    oMu, oCov, olam, otr = orc.em_iterations(X, Mu0, Cov0, lam0, mp_, 1)
    print(f"d={d} K={K} N={X.shape[0]}  CUDA vs exact: sumLL {abs(em.sumLL - float(mLL)) / abs(float(mLL)):.1e} Mu {rel(em.Mu, eMu):.1e} Cov {rel(em.Cov, eCov):.1e} lambda {np.abs(em.lambda_ - eLam).max():.1e}"
          f"   | oracle vs exact: sumLL {abs(otr[0] - float(mLL)) / abs(float(mLL)):.1e} Mu {rel(oMu, eMu):.1e} Cov {rel(oCov, eCov):.1e} lambda {np.abs(olam - eLam).max():.1e}   [{em.kernel_plan()[:12]}]")
    em.finalize()
