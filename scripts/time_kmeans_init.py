"""Times the first step of a fresh object (k-means initialiser: D^2 seeding + 15 Lloyd steps) at the headline shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlfortran_b200 import mlf_algo_emgmm, synth
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
X, centres, _ = synth.make_mixture_device(16, 64, N, 0, N, 8.0, 1.0, 3, dev)
for attempt in ("first object of the process (includes one-time kernel loading)", "second object"):
    em = mlf_algo_emgmm()
    assert em.init(X, 64, 1.0) == 0
    em.set_seed(7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info, dt = em.step(1)
    t1 = time.perf_counter()
    assert info == 0, em.last_error()
    d = np.linalg.norm(em.Mu[:, None, :] - centres[None], axis=2)
    print({"N": N, "init_seconds": t1 - t0, "centres_recovered_within_1": int((d.min(axis=0) < 1.0).sum()), "of": 64, "which": attempt})
    em.finalize()
