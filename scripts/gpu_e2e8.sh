#!/bin/bash
# 8-GPU e2e: the streamed step against the plain-copy ceiling, page-locked vs write-combined host buffers
G=${G:-8}
for hb in ${HBS:-pinned wc}; do
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --steps 10 --warmup 3 --no-cpu --host-buffer $hb \
    > gpurun_out/bench_e2e${G}_$hb.json 2> gpurun_out/bench_e2e${G}_$hb.err; echo "bench $G x $hb exit $?"
done
nvidia-smi topo -m > gpurun_out/topo_${G}gpu.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/topo_${G}gpu.txt
python - <<'PY'
import json, os
G=os.environ.get("G","8")
for hb in os.environ.get("HBS", "pinned wc").split():
    try:
        j=json.loads(open(f"gpurun_out/bench_e2e{G}_{hb}.json").read().strip().splitlines()[-1]); r=j["roofline"]
        print(hb, "n_gpus", j["n_gpus"], "ms/step %.2f iter/s %.3f frac %.4f"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"]))
        print("  e2e", j["e2e"] and {k: j["e2e"][k] for k in ("value","ms_per_step","h2d_gbs","h2d_probe_gbs","host_buffer","host_affinity")})
    except Exception as e:
        print(hb, "ERR", e, open(f"gpurun_out/bench_e2e{G}_{hb}.err").read()[-800:])
PY
