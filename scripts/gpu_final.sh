#!/bin/bash
# the round's final record on one GPU, all with the same build
T=${TAG:-r02m}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$T.log; tail -3 gpurun_out/pytest_gpu_$T.log
python __graft_entry__.py smoke > gpurun_out/smoke_$T.log 2>&1; tail -1 gpurun_out/smoke_$T.log
python bench.py > gpurun_out/bench_${T}_1gpu_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python bench.py --mode infer > gpurun_out/infer_$T.json 2> gpurun_out/infer.err; echo "infer exit $?"
python bench.py --points 2e7 --dims 128 --comps 32 --sigma-c 4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${T}_cfg4_d128.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"
python bench.py --points 1.25e8 --dims 8 --comps 256 --cov-type diag --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${T}_cfg5_shard.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 exit $?"
python bench.py --points 1e6 --dims 2 --comps 8 --sigma-c 6 --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_${T}_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit $?"
MLF_B200_DEBUG=1 python scripts/time_kmeans_init.py 1e8 2>&1 | grep -E "k-means|init_seconds" > gpurun_out/kmeans_init_$T.txt; cat gpurun_out/kmeans_init_$T.txt
SCS="8 4 2 1 0.5" bash scripts/gpu_overlap.sh > gpurun_out/overlap_$T.txt 2>&1; cp gpurun_out/overlap.jsonl gpurun_out/overlap_$T.jsonl; cat gpurun_out/overlap_$T.txt
B="python bench.py --points 4e6 --steps 2 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_prof.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}_n4e6.csv $B > gpurun_out/ncu_list.log 2>&1
B2="python bench.py --points 4e6 --steps 4 --warmup 3 --no-e2e --no-cpu --sigma-c 1"
$B2 > gpurun_out/plain_prof_sc1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_scatter16' -s 1 -c 1 -o gpurun_out/prof_${T}_scatter16 $B2 > gpurun_out/ncu_full_sc.log 2>&1
tail -2 gpurun_out/ncu_full_sc.log
python - <<'PY'
import json, os
T=os.environ.get("TAG","r02m")
def last(f):
    return json.loads(open(f).read().strip().splitlines()[-1])
j=last(f"gpurun_out/bench_{T}_1gpu_default.json"); r=j["roofline"]
print("default: ms/step %.2f iter/s %.3f frac %.4f"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"]), "e2e", {k: j["e2e"][k] for k in ("ms_per_step","h2d_gbs","h2d_probe_gbs")}, "cpu", j["cpu_baseline"]["value"], "parity", j["parity_check"]["pass"], "clocks", j["clocks"])
for n in ("cfg4_d128","cfg5_shard","cfg2"):
    j=last(f"gpurun_out/bench_{T}_{n}.json"); r=j["roofline"]
    print(n, "ms/step %.3f iter/s %.3f frac %.4f"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"]), {k: round(v,3) for k,v in r["per_iteration_ms"].items()})
print(open(f"gpurun_out/infer_{T}.json").read()[-700:])
PY
