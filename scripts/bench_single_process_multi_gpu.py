"""Single-process multi-GPU object (mlf_b200_emgmm_create_multi): N=1e8, d=16, K=64 over every visible GPU of the box, one
host process, library-internal peer-memory all-reduce.  Prints one JSON line (wall-clock timing of mlf_step)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlfortran_b200 import mlf_algo_emgmm, synth

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
G = torch.cuda.device_count()
d, K = 16, 64
dev0 = torch.device("cuda", 0)
Xh = torch.empty((N, d), dtype=torch.float64, pin_memory=True)
centres = None
chunk = N // 4
for c in range(4):   # generate on GPU 0 in quarters (keeps GPU 0's memory free for its shard afterwards)
    Xg, centres, _ = synth.make_mixture_torch(d, K, chunk, 8.0, 1.0, 3, dev0, stream_seed=3 + c)
    Xh[c * chunk:(c + 1) * chunk].copy_(Xg)
    del Xg
torch.cuda.synchronize(); torch.cuda.empty_cache()
Mu0, Cov0, lam0 = synth.injected_init(centres, 3)
em = mlf_algo_emgmm()
t0 = time.perf_counter()
assert em.init(Xh.numpy(), K, 1.0, devices=list(range(G))) == 0
em.inject(Mu0, Cov0, lam0)
info, _ = em.step(3)
assert info == 0, em.last_error()
t1 = time.perf_counter()
steps = 10
info, dt = em.step(steps)
assert info == 0, em.last_error()
print(json.dumps({"mode": "single process, %d GPUs, peer-memory all-reduce" % G, "N": N, "d": d, "K": K, "plan": em.kernel_plan(),
                  "init_plus_3_iterations_s": t1 - t0, "ms_per_iteration": dt / steps * 1e3, "em_iter_per_s": steps / dt,
                  "point_comp_per_s": N * K * steps / dt, "sumLL": em.sumLL}))
