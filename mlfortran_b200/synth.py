"""Synthetic Gaussian-mixture workloads in the shape of the reference's own test program.

Recipe follows tests/test_emgmm.f90:105-124 (testGMM): centres XC ~ sigmaC*N(0,I); a common
spectrum w_i = sqrt(sigmaX*(1/i)/mean_j(1/j)) (:116-117); per component C12 = O*diag(w) with O a
Haar-random orthogonal matrix (src/mlf_matrix.f90:149-180, Stewart 1980) and NX points
x = c_k + C12*N(0,I) (:118-121); the points are then shuffled (:122-124).  The reference draws from
an unseeded RANDOM_NUMBER stream, so only the distribution -- not the bits -- can be reproduced; our
streams are seeded NumPy / torch generators.

Layouts: X (N, d) C-contiguous == Fortran X(d,N); Mu (K, d); Cov (K, d, d); lambda (K,).
"""
from __future__ import annotations

import numpy as np


def spectrum(d: int, sigma_x: float) -> np.ndarray:
    w = 1.0 / np.arange(1, d + 1, dtype=np.float64)
    return np.sqrt(sigma_x * w / w.mean())


def haar_orthogonal(rng: np.random.Generator, d: int) -> np.ndarray:
    q, r = np.linalg.qr(rng.standard_normal((d, d)))
    return q * np.sign(np.diag(r))


def make_mixture(d: int, K: int, n_per_comp: int, sigma_c: float, sigma_x: float, seed: int, shuffle: bool = True):
    """Host (NumPy) generator; returns dict(X, centres, C12, labels) with labels 1-based."""
    rng = np.random.default_rng(seed)
    XC = sigma_c * rng.standard_normal((K, d))
    w = spectrum(d, sigma_x)
    C12 = np.empty((K, d, d))
    X = np.empty((K * n_per_comp, d))
    lab = np.empty(K * n_per_comp, dtype=np.int32)
    for k in range(K):
        C12[k] = haar_orthogonal(rng, d) * w  # O*diag(w): column i scaled by w_i
        X[k * n_per_comp:(k + 1) * n_per_comp] = XC[k] + rng.standard_normal((n_per_comp, d)) @ C12[k].T
        lab[k * n_per_comp:(k + 1) * n_per_comp] = k + 1
    if shuffle:
        perm = rng.permutation(K * n_per_comp)
        X, lab = X[perm], lab[perm]
    return {"X": np.ascontiguousarray(X), "centres": XC, "C12": C12, "labels": lab}


def injected_init(centres: np.ndarray, seed: int, jitter: float = 0.5):
    """What stepF sets after its k-means call (src/unsupervised/mlf_emgmm.f90:140-145): Mu near the
    clusters, Cov = I, lambda = 1/K.  The reference's k-means is RNG-driven, so tests inject this."""
    rng = np.random.default_rng(seed + 7919)
    K, d = centres.shape
    Mu = centres + jitter * rng.standard_normal((K, d))
    Cov = np.tile(np.eye(d), (K, 1, 1))
    lam = np.full(K, 1.0 / K)
    return Mu, Cov, lam


def make_mixture_torch(d: int, K: int, N: int, sigma_c: float, sigma_x: float, seed: int, device, chunk: int = 1 << 22,
                       stream_seed: int | None = None):
    """Device generator for the large bench shapes (same distribution; component of point i is
    drawn uniformly, which is what shuffling equal-sized blocks yields in the limit).
    Returns (X [N,d] float64 on device, centres [K,d] numpy, C12 [K,d,d] numpy)."""
    import torch

    rng = np.random.default_rng(seed)
    XC = sigma_c * rng.standard_normal((K, d))
    w = spectrum(d, sigma_x)
    C12 = np.stack([haar_orthogonal(rng, d) * w for _ in range(K)])
    g = torch.Generator(device=device)
    g.manual_seed(seed if stream_seed is None else stream_seed)  # mixture from `seed`, sample stream per rank
    X = torch.empty((N, d), dtype=torch.float64, device=device)
    tXC = torch.as_tensor(XC, device=device)
    tC12T = torch.as_tensor(np.ascontiguousarray(C12.transpose(0, 2, 1)), device=device)
    for s in range(0, N, chunk):
        n = min(chunk, N - s)
        comp = torch.randint(0, K, (n,), generator=g, device=device)
        z = torch.randn((n, d), generator=g, dtype=torch.float64, device=device)
        # x = c_k + C12_k z, batched by sorting into components would cost a sort; use bmm on gathered factors in slices
        out = X[s:s + n]
        step = 1 << 18
        for t in range(0, n, step):
            cc = comp[t:t + step]
            out[t:t + step] = tXC[cc] + torch.bmm(z[t:t + step].unsqueeze(1), tC12T[cc]).squeeze(1)
    return X, XC, C12


def make_mixture_device(d: int, K: int, n_local: int, n0: int, n_total: int, sigma_c: float, sigma_x: float, seed: int, device):
    """The library's own device generator (`mlf_b200_synth_gmm`, csrc/mlf_synth.cu): points [n0, n0 + n_local) of the
    n_total-point job written by one CUDA kernel into a torch-allocated HBM buffer (torch is only the allocator here).
    Every value is a pure function of (seed, global point index), so the shards of any number of ranks form the same data set.
    Returns (X [n_local, d] float64 on device, centres [K, d] numpy, C12 [K, d, d] numpy with C12[k] = O_k diag(w))."""
    import torch

    from . import _lib

    lib = _lib.load()
    dev = torch.device(device)
    X = torch.empty((n_local, d), dtype=torch.float64, device=dev)
    centres = np.empty((K, d))
    C12f = np.empty((K, d, d))  # Fortran C12(a, b, k): C12f[k, b, a]
    torch.cuda.synchronize(dev)
    rc = lib.mlf_b200_synth_gmm(X.data_ptr(), n0, n_local, n_total, d, K, float(sigma_c), float(sigma_x), int(seed),
                                dev.index if dev.index is not None else torch.cuda.current_device(), 1,  # MLF_B200_X_ON_DEVICE
                                centres.ctypes.data, C12f.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"mlf_b200_synth_gmm failed with {rc}")
    return X, centres, np.ascontiguousarray(C12f.transpose(0, 2, 1))


def mixture_params(d: int, K: int, sigma_c: float, sigma_x: float, seed: int):
    """centres and C12 of the device generator's mixture (host part of `mlf_b200_synth_gmm`, no GPU needed)."""
    from . import _lib

    centres = np.empty((K, d))
    C12f = np.empty((K, d, d))
    rc = _lib.load().mlf_b200_synth_gmm(None, 0, 0, 1, d, K, float(sigma_c), float(sigma_x), int(seed), 0, 0,
                                        centres.ctypes.data, C12f.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"mlf_b200_synth_gmm failed with {rc}")
    return centres, np.ascontiguousarray(C12f.transpose(0, 2, 1))
