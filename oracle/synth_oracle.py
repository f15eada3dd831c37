"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the device workload generator (mlfortran_b200/csrc/mlf_synth.cu).

Only tests/ may import this.  It restates, independently and vectorised, what `mlf_b200_synth_gmm` defines:
the recipe of the reference's own EM test program (tests/test_emgmm.f90:105-124: RandN centres, `randOrthogonal(C12, weight)`
of src/mlf_matrix.f90:153-180, `RandN(X, X0, C12)` of src/mlf_rand.f90:241-257, `randPerm`) driven by a counter-based stream
(Philox4x32-10 + Box-Muller, keyed Feistel permutation) instead of gfortran's unseeded RANDOM_NUMBER.  The reference's bits
cannot be reproduced (parity is distributional); the CUDA generator must match THIS file to rounding (libm vs CUDA libm).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
U64 = np.uint64
STREAM_POINTS, STREAM_CENTRES, STREAM_ORTHO, STREAM_PERM = 0, 1, 2, 3
_M0, _M1, _W0, _W1 = U64(0xD2511F53), U64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
_MASK32 = U64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al., SC'11); counters are broadcastable integer arrays, key two Python ints."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=U64) & _MASK32 for c in (c0, c1, c2, c3)])
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        h0, l0, h1, l1 = p0 >> U64(32), p0 & _MASK32, p1 >> U64(32), p1 & _MASK32
        c0, c1, c2, c3 = h1 ^ c1 ^ U64(k0), l1, h0 ^ c3 ^ U64(k1), l0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u01(hi, lo):
    a = ((hi << U64(32)) | lo) >> U64(11)
    return (a.astype(np.float64) + 0.5) * 2.0 ** -53


def normal_pairs(seed: int, stream: int, idx, pair):
    """Box-Muller pair (z0, z1) number `pair` of item `idx` in `stream`."""
    idx = np.asarray(idx, dtype=U64)
    r0, r1, r2, r3 = philox4x32_10(idx & _MASK32, idx >> U64(32), pair, stream, seed & 0xFFFFFFFF, seed >> 32)
    u1, u2 = _u01(r0, r1), _u01(r2, r3)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 6.283185307179586476925286766559 * u2
    return rad * np.cos(ang), rad * np.sin(ang)


def _fmix32(x):
    x = x & _MASK32
    x ^= x >> U64(16)
    x = (x * U64(0x85EBCA6B)) & _MASK32
    x ^= x >> U64(13)
    x = (x * U64(0xC2B2AE35)) & _MASK32
    x ^= x >> U64(16)
    return x


def permutation(seed: int, n: int, idx):
    """pi(idx): keyed bijection of [0, n) (4-round balanced Feistel on an even number of bits, cycle-walked)."""
    bits = 2
    while bits < 62 and (1 << bits) < n:
        bits += 2
    half = U64(bits // 2)
    mask = U64((1 << (bits // 2)) - 1)
    rk = [U64(int(v)) for v in philox4x32_10(0, 0, 0, STREAM_PERM, seed & 0xFFFFFFFF, seed >> 32)]
    out = np.array(idx, dtype=U64, copy=True).reshape(-1)
    todo = np.ones(out.shape, dtype=bool)
    while todo.any():
        i = out[todo]
        L, R = i >> half, i & mask
        for r in range(4):
            F = _fmix32(((R & _MASK32) * U64(0x9E3779B1) + rk[r]) & _MASK32) & mask
            L, R = R, L ^ F
        out[todo] = (L << half) | R
        todo = out >= U64(n)
    return out.reshape(np.shape(idx))


def spectrum(d: int, sigma_x: float) -> np.ndarray:
    w = 1.0 / np.arange(1, d + 1, dtype=np.float64)
    return np.sqrt(sigma_x * w / (w.sum() / d))


def rand_orthogonal(seed: int, item: int, n: int, sigma=None) -> np.ndarray:
    """randOrthogonal(H, Sigma) (src/mlf_matrix.f90:153-180): returns H (n x n, math indexing H[row, col])."""
    H = np.eye(n)
    D = np.ones(n)
    pair = 0
    for k in range(n - 1):
        m = n - k
        npairs = (m + 1) // 2
        z0, z1 = normal_pairs(seed, STREAM_ORTHO, item, np.arange(pair, pair + npairs))
        pair += npairs
        v = np.empty(2 * npairs)
        v[0::2], v[1::2] = z0, z1
        v = v[:m].copy()
        nrm = 0.0
        for x in v:          # sequential sum, as the C++ loop
            nrm += x * x
        D[k] = -1.0 if np.signbit(v[0]) else 1.0
        v[0] -= D[k] * np.sqrt(nrm)
        vv = 0.0
        for x in v:
            vv += x * x
        w = np.array([2.0 * _seqdot(v, H[k:, c]) / vv for c in range(n)])
        H[k:, :] -= np.outer(v, w)
    D[n - 1] = np.prod(D[: n - 1]) * (1.0 if n % 2 else -1.0)
    if sigma is not None:
        D = D * sigma
    return H * D[None, :]


def _seqdot(a, b):
    s = 0.0
    for x, y in zip(a, b):
        s += x * y
    return s


def mixture_params(d: int, K: int, sigma_c: float, sigma_x: float, seed: int):
    """centres [K, d] and C12 [K, d, d] (math indexing C12[k, row, col])."""
    w = spectrum(d, sigma_x)
    hd = (d + 1) // 2
    centres = np.empty((K, 2 * hd))
    z0, z1 = normal_pairs(seed, STREAM_CENTRES, np.arange(K)[:, None], np.arange(hd)[None, :])
    centres[:, 0::2], centres[:, 1::2] = z0, z1
    centres = sigma_c * centres[:, :d]
    C12 = np.stack([rand_orthogonal(seed, k, d, w) for k in range(K)])
    return np.ascontiguousarray(centres), C12


def points(d: int, K: int, n0: int, nX: int, n_total: int, sigma_c: float, sigma_x: float, seed: int):
    """Points [n0, n0 + nX) of the n_total-point job: X [nX, d], labels (1-based component of each point)."""
    centres, C12 = mixture_params(d, K, sigma_c, sigma_x, seed)
    src = permutation(seed, n_total, np.arange(n0, n0 + nX, dtype=np.uint64))
    hd = (d + 1) // 2
    z = np.empty((nX, 2 * hd))
    z0, z1 = normal_pairs(seed, STREAM_POINTS, src[:, None], np.arange(hd)[None, :])
    z[:, 0::2], z[:, 1::2] = z0, z1
    comp = (src % np.uint64(K)).astype(np.int64)
    X = np.empty((nX, d))
    for i in range(nX):       # fma chain of the kernel in plain arithmetic: agreement to rounding, not to the bit
        X[i] = centres[comp[i]] + C12[comp[i]] @ z[i, :d]
    return X, (comp + 1).astype(np.int32), centres, C12
