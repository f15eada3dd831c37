"""Independent NumPy restatement of the reference EM-GMM iteration -- TEST INFRASTRUCTURE.

Written separately from oracle/emgmm_oracle.c (vectorised, numpy.linalg.eigh instead of
dsytrd/dorgtr/dsteqr) so that the two restatements check each other.  PARITY UNPINNED: see
emgmm_oracle.c.  Layout conventions as in emgmm_oracle.py (X (N,d), Mu (K,d), Cov (K,d,d), P (K,N)).
"""
from __future__ import annotations

import numpy as np

VALM_FACTOR = float(np.float32(10e-9))  # src/mlf_matrix.f90:283 -- default-real literal


def inverse_sym_matrix(Cm):
    """src/mlf_matrix.f90:229-300 with doEigen=.TRUE. (eigen branch :256-298)."""
    d = Cm.shape[0]
    w, V = np.linalg.eigh(Cm)
    valM = VALM_FACTOR * w.max()
    with np.errstate(divide="ignore"):
        f = np.where(w <= valM, 1.0 / valM, 1.0 / w)
    lnDet = 0.5 * (d * np.log(2.0 * np.arccos(-1.0)) + np.sum(np.log(f)))
    return (V * f) @ V.T, lnDet


def exp_step(X, Mu, Cov, lam, cov_diag=False):
    """src/unsupervised/mlf_emgmm.f90:191-207 + mlf_gaussian.f90:63-79.  cov_diag: extension (diagonal of Cov only)."""
    N, d = X.shape
    K = Mu.shape[0]
    P = np.empty((K, N))
    sumLL = 0.0
    for k in range(K):
        invC, lnDet = inverse_sym_matrix(np.diag(np.diag(Cov[k])) if cov_diag else Cov[k])
        Z = X - Mu[k]
        v = -0.5 * np.einsum("ia,ab,ib->i", Z, invC, Z)
        with np.errstate(divide="ignore"):
            sumLL += N * (np.log(lam[k]) - lnDet) + v.sum()
        P[k] = lam[k] * np.exp(v - lnDet)
    with np.errstate(invalid="ignore", divide="ignore"):
        P /= P.sum(axis=0, keepdims=True)
    return P, sumLL


def max_step(X, P, min_proba=0.0, cov_diag=False):
    """src/unsupervised/mlf_emgmm.f90:174-188 + mlf_gaussian.f90:41-51 + mlf_matrix.f90:99-147."""
    N, d = X.shape
    K = P.shape[0]
    Mu = np.empty((K, d))
    Cov = np.empty((K, d, d))
    lam = np.empty(K)
    minW = 1e-3 / float(np.float32(N))
    for k in range(K):
        s = P[k].sum()
        W = P[k] / s
        ms = 1.0 / (1.0 - np.sum(W * W))
        Mu[k] = W @ X
        keep = ~(W < minW)
        B = (ms * np.sqrt(W[keep]))[:, None] * (X[keep] - Mu[k])
        Cov[k] = B.T @ B
        if cov_diag:
            Cov[k] = np.diag(np.diag(Cov[k]))
        lam[k] = s
    if min_proba > 0:
        lam = np.maximum(lam, min_proba * lam.max())
    return Mu, Cov, lam / lam.sum()


def em_iterations(X, Mu, Cov, lam, min_proba, niter, cov_diag=False):
    trace = []
    P = None
    for _ in range(niter):
        P, ll = exp_step(X, Mu, Cov, lam, cov_diag)
        trace.append(ll)
        Mu, Cov, lam = max_step(X, P, min_proba, cov_diag)
    return Mu, Cov, lam, np.array(trace), P


def get_class(P):
    """src/mlf_models.f90:248 -- MAXLOC(Proba, DIM=2): first maximum, 1-based."""
    return (np.argmax(P, axis=0) + 1).astype(np.int32)
