#!/bin/bash
# the threshold protocol's worst case at the headline shape -- overlapping mixtures (sigmaC small) in both
# minProba regimes.  One line per run into gpurun_out/overlap.jsonl.
N=${N:-1e8}
: > gpurun_out/overlap.jsonl
for mp in 1 0; do for sc in ${SCS:-8 4 2 1 0.5}; do
  timeout -s KILL ${BTO:-150} python bench.py --points $N --steps ${STEPS:-10} --warmup 3 --no-e2e --no-cpu --sigma-c $sc --min-proba $mp > gpurun_out/ov_tmp.json 2> gpurun_out/ov_tmp.err \
    && tail -1 gpurun_out/ov_tmp.json >> gpurun_out/overlap.jsonl || { echo "{\"failed\": \"sigmaC=$sc minProba=$mp\"}" >> gpurun_out/overlap.jsonl; tail -5 gpurun_out/ov_tmp.err; }
done; done
python - <<'PY'
import json
for ln in open("gpurun_out/overlap.jsonl"):
    j=json.loads(ln)
    if "failed" in j: print(j); continue
    r=j["roofline"]; c=j["config"]
    print("sigmaC %-4s minProba %s  ms/iter %8.2f  sweeps %.2f  side-list pairs/iter %10.1f  kept pairs/point %6.3f  frac %.3f  sumLL %.6e"%(
        c["sigmaC"], c["minProba"], j["ms_per_step"], r["sweeps_per_iteration"], r["side_list_pairs"]/j["steps"], r["kept_pairs_per_point"] or 0, r["frac"], j["sumLL_last"]))
PY
