#!/bin/bash
# default bench (all legs) with the page-locked and the write-combined host buffer, then the reference arm
for hb in ${HBS:-pinned wc}; do
  python bench.py --host-buffer $hb > gpurun_out/bench_full_$hb.json 2> gpurun_out/bench_full_$hb.err; echo "bench $hb exit $?"
done
python - <<'PY'
import json, os
for hb in os.environ.get("HBS", "pinned wc").split():
    try:
        j=json.loads(open(f"gpurun_out/bench_full_{hb}.json").read().strip().splitlines()[-1]); r=j["roofline"]
        print(hb, "ms/step %.2f iter/s %.3f frac %.4f"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"]), r["logits"])
        print("  e2e", j["e2e"] and {k: j["e2e"][k] for k in ("value","ms_per_step","h2d_gbs","h2d_probe_gbs","host_buffer")})
        print("  cpu", j["cpu_baseline"] and j["cpu_baseline"]["value"])
        print("  parity", j["parity_check"] and {k: v for k, v in j["parity_check"].items() if k not in ("sample","tolerances")})
        print("  clocks", j["clocks"], "launches", j["gpu_launches"])
    except Exception as e:
        print(hb, "ERR", e, open(f"gpurun_out/bench_full_{hb}.err").read()[-800:])
PY
