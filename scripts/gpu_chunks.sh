#!/bin/bash
# e2e step against the plain-copy probe for different numbers of upload pieces (MLF_B200_STREAM_CHUNKS) on G GPUs
G=${G:-4}
for c in ${CHUNKS:-16 4 2 64}; do
  MLF_B200_STREAM_CHUNKS=$c timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $G --steps 3 --warmup 3 --e2e-steps 6 --no-cpu \
    > gpurun_out/bench_chunks_${G}gpu_$c.json 2> gpurun_out/bench_chunks_${G}gpu_$c.err
  python - $G $c <<'PY'
import json, sys
G, c = sys.argv[1], sys.argv[2]
try:
    j=json.loads(open(f"gpurun_out/bench_chunks_{G}gpu_{c}.json").read().strip().splitlines()[-1])
    e=j["e2e"]; print("GPUs", G, "chunks", c, "e2e ms/step %.2f  h2d %.2f GB/s per GPU  probe %.2f GB/s"%(e["ms_per_step"], e["h2d_gbs"], e["h2d_probe_gbs"]))
except Exception as ex:
    print("ERR", ex, open(f"gpurun_out/bench_chunks_{G}gpu_{c}.err").read()[-600:])
PY
done
