"""TEST INFRASTRUCTURE ONLY -- 50-digit (mpmath) evaluation of the reference EM-GMM iteration for tiny shapes.

A third restatement of SURVEY Appendix A (src/unsupervised/mlf_emgmm.f90:174-207, src/unsupervised/mlf_gaussian.f90:41-79,
src/mlf_matrix.f90:99-147,229-300) in arbitrary precision: it takes the fp64 inputs exactly and rounds nothing on the way, so
the distance between it and the fp64 oracle is the oracle's own rounding error.  tests/test_oracle_kat.py uses it to show that
error is orders of magnitude inside the 1e-10 / 1e-9 parity budget the CUDA path is held to.  Only tests/ may import this."""
from __future__ import annotations

import mpmath as mp
import numpy as np

mp.mp.dps = 50
VALM_FACTOR = mp.mpf(float(np.float32(10e-9)))  # src/mlf_matrix.f90:283 -- default-real literal, taken exactly


def _inverse_sym_matrix(C):
    d = C.rows
    E, Q = mp.eigsy(C)
    e = [E[i] for i in range(d)]
    valM = VALM_FACTOR * max(e)
    f = [1 / valM if e[i] <= valM else 1 / e[i] for i in range(d)]
    lnDet = (d * mp.log(2 * mp.acos(-1)) + sum(mp.log(x) for x in f)) / 2   # the reference's sign
    invC = Q * mp.diag(f) * Q.T
    return invC, lnDet


def em_iteration(X, Mu, Cov, lam, min_proba):
    """One ExpStep + MaxStep.  Inputs: numpy fp64 arrays X (N,d), Mu (K,d), Cov (K,d,d), lam (K,).  Returns mp values."""
    N, d = X.shape
    K = Mu.shape[0]
    Xm = [[mp.mpf(float(v)) for v in row] for row in X]
    P = [[None] * N for _ in range(K)]
    sumLL = mp.mpf(0)
    for k in range(K):
        invC, lnDet = _inverse_sym_matrix(mp.matrix([[mp.mpf(float(v)) for v in row] for row in Cov[k]]))
        mu = [mp.mpf(float(v)) for v in Mu[k]]
        lk = mp.mpf(float(lam[k]))
        LL = N * (mp.log(lk) - lnDet)
        for i in range(N):
            z = [Xm[i][a] - mu[a] for a in range(d)]
            v = -sum(z[a] * invC[a, b] * z[b] for a in range(d) for b in range(d)) / 2
            LL += v
            P[k][i] = lk * mp.exp(v - lnDet)
        sumLL += LL
    for i in range(N):
        s = sum(P[k][i] for k in range(K))
        for k in range(K):
            P[k][i] /= s
    minW = mp.mpf(1e-3) / mp.mpf(float(np.float32(N)))   # 1d-3 / real(NX): both taken exactly
    newMu, newCov, newLam = [], [], []
    for k in range(K):
        s = sum(P[k])
        W = [p / s for p in P[k]]
        ms = 1 / (1 - sum(w * w for w in W))
        mu = [sum(W[i] * Xm[i][a] for i in range(N)) for a in range(d)]
        C = [[mp.mpf(0)] * d for _ in range(d)]
        for i in range(N):
            if W[i] < minW:
                continue
            b = [ms * mp.sqrt(W[i]) * (Xm[i][a] - mu[a]) for a in range(d)]
            for a in range(d):
                for c in range(d):
                    C[a][c] += b[a] * b[c]
        newMu.append(mu)
        newCov.append(C)
        newLam.append(s)
    if min_proba > 0:
        top = max(newLam)
        newLam = [max(l, mp.mpf(float(min_proba)) * top) for l in newLam]
    tot = sum(newLam)
    newLam = [l / tot for l in newLam]
    return newMu, newCov, newLam, sumLL, P
