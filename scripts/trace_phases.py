"""Debug builds only (make -C mlfortran_b200/csrc EXTRA=-DMLF_TRACE=1 after touching emgmm_fused16.cu): dumps the phase timeline
of block 0's sixteen warps (clock64 at the start of P1, P2, the group barrier, P3a, P3b and the tile's end) for 64 tiles of the last
sweep.  Usage (GPU box): python scripts/trace_phases.py [N] [sigma_c] -> gpurun_out/trace_phases.npy"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlfortran_b200 import mlf_algo_emgmm, synth, _lib
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
sc = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
dev = torch.device("cuda", 0)
X, centres, _ = synth.make_mixture_device(16, 64, N, 0, N, sc, 1.0, 3, dev)
Mu0, Cov0, lam0 = synth.injected_init(centres, 3)
em = mlf_algo_emgmm()
assert em.init(X, 64, 1.0) == 0
em.inject(Mu0, Cov0, lam0)
assert em.step(6)[0] == 0
lib = C.CDLL(_lib.LIB_PATH)
out = np.zeros(64 * 16 * 8, dtype=np.int64)
assert lib.mlf_b200_trace_read(out.ctypes.data_as(C.c_void_p), out.size) == 0
np.save("gpurun_out/trace_phases.npy", out.reshape(64, 16, 8))
t = out.reshape(64, 16, 8)[:, :, :6].astype(np.float64)
d = np.diff(t, axis=2)
names = ["P1", "P2", "barrier", "P3a", "P3b"]
print("mean clocks per phase (all warps, tiles 8..71):", {n: float(d[:, :, i].mean()) for i, n in enumerate(names)})
print("tile period per warp:", float(np.diff(t[:, :, 0], axis=0).mean()))
