"""Config 3 at a tenth of its size: N = 1e7, d = 16, K = 64, five EM iterations of the CUDA path (through the C-ABI) against the
oracle on the same points, with the quadratic-form logits (default) and with whitening, on the bench mixture (sigmaC = 8) and on
a moderately overlapping one (sigmaC = 2).  One-off evidence run (the oracle needs ~20 s per case on 16 cores); prints one JSON
line per case.  Usage (GPU box): python scripts/dev_parity_1e7.py > gpurun_out/parity_1e7.jsonl"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mlfortran_b200 import mlf_algo_emgmm, synth
from oracle.emgmm_oracle import get

orc = get()
d, K, npc, iters = 16, 64, 156250, 5
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
for sc in (8.0, 2.0):
    data = synth.make_mixture(d, K, npc, sc, 1.0, 3)
    X = data["X"]
    Mu0, Cov0, lam0 = synth.injected_init(data["centres"], 3)
    t0 = time.time()
    oMu, oCov, olam, otr = orc.em_iterations(X, Mu0, Cov0, lam0, 1.0, iters)
    t_or = time.time() - t0
    for mode in ("1", "0"):
        os.environ["MLF_B200_QF"] = mode
        em = mlf_algo_emgmm()
        assert em.init(X, K, 1.0) == 0
        em.inject(Mu0, Cov0, lam0)
        em.set_profile(True)
        info, dt = em.step(iters)
        assert info == 0, em.last_error()
        tr = em.sumLL_trace()
        print(json.dumps({"N": X.shape[0], "d": d, "K": K, "sigmaC": sc, "iterations": iters, "logits": "quadratic form (guarded)" if mode == "1" else "whitening",
                          "qf_sweeps": em.qf_stats()["qf_sweeps"], "sweeps": em.counters()["sweeps"],
                          "sumLL_rel_max_over_iterations": float(np.abs((tr - otr) / otr).max()), "Mu_rel": rel(em.Mu, oMu), "Cov_rel": rel(em.Cov, oCov),
                          "lambda_abs": float(np.abs(em.lambda_ - olam).max()), "oracle_seconds": round(t_or, 1), "gpu_seconds": round(dt, 3)}), flush=True)
        em.finalize()
