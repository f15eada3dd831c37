"""Multi-rank host logic on CPU (gloo, world_size 2): the shard partition, the all-reduce callback the engine calls,
and the two-pass sufficient-statistics protocol the engine runs across ranks (pass-A stats -> all-reduce -> global
mean/threshold -> pass-B scatter -> all-reduce -> finalize) checked against the oracle's whole-data iteration.
The per-shard arithmetic is a NumPy stand-in (the product has no CPU compute path); what is under test is that the
*decomposition* -- global N in minW, global s_k in the threshold, summed statistics -- reproduces the reference."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from mlfortran_b200 import parallel, synth


def test_shard_bounds_partition():
    for n in [0, 1, 7, 8, 1000, 10 ** 8 + 3]:
        for w in [1, 2, 3, 8]:
            b = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_stats_A(X, P):
    # what k_estep accumulates per component: [sum g, sum g^2, sum g x]
    return np.concatenate([P.sum(1, keepdims=True), (P * P).sum(1, keepdims=True), P @ X], axis=1)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from oracle import emgmm_numpy as onp

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ar = parallel.make_allreduce(None, None)

        def allreduce(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert ar(a.ctypes.data, a.size, 0) == 0  # same (pointer, count, stream) signature the C engine uses
            return a

        # 1. callback semantics on raw host memory
        buf = np.arange(5, dtype=np.float64) + 10 * rank
        allreduce(buf)
        exp = sum(np.arange(5, dtype=np.float64) + 10 * r for r in range(world))
        ok1 = bool(np.array_equal(buf, exp))

        # 2. sharded EM iteration == whole-data reference iteration
        d, K = 3, 4
        data = synth.make_mixture(d, K, 250, 5.0, 1.0, 11)
        X = data["X"]
        Mu, Cov, lam = synth.injected_init(data["centres"], 11)
        lo, hi = parallel.shard_bounds(X.shape[0], rank, world)
        Xl = X[lo:hi]
        for min_proba in (0.0, 1.0):
            Mu_w, Cov_w, lam_w, tr_w, _ = onp.em_iterations(X, Mu, Cov, lam, min_proba, 3)  # whole data
            m, c, l = Mu.copy(), Cov.copy(), lam.copy()
            for _ in range(3):
                P, _ = onp.exp_step(Xl, m, c, l)                   # responsibilities are per-point: shard-local
                nglob = allreduce(np.array([float(Xl.shape[0])]))[0]
                A = allreduce(_shard_stats_A(Xl, P))                 # all-reduce #1: K x (2+d)
                s = A[:, 0]
                m = A[:, 2:] / s[:, None]
                thr = (1e-3 / float(np.float32(nglob))) * s          # global N, global s_k
                B = np.zeros((K, d, d))
                for k in range(K):
                    keep = ~(P[k] < thr[k])
                    Z = Xl[keep] - m[k]
                    B[k] = (Z.T * P[k][keep]) @ Z
                B = allreduce(B)                                     # all-reduce #2: K x d x d
                ms = 1.0 / (1.0 - A[:, 1] / (s * s))
                c = (ms * ms / s)[:, None, None] * B
                l = np.maximum(s, min_proba * s.max()) if min_proba > 0 else s
                l = l / l.sum()
            ok2 = (np.abs(m - Mu_w).max() < 1e-11 and np.abs(c - Cov_w).max() < 1e-11 and np.abs(l - lam_w).max() < 1e-13)
            if not ok2:
                break
        q.put((rank, ok1, bool(ok2)))
    finally:
        dist.destroy_process_group()


def test_two_rank_protocol_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "all-reduce callback did not sum across ranks"
    assert all(r[2] for r in res), "sharded two-pass protocol deviates from the whole-data iteration"
