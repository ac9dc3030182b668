#!/bin/bash
# The other BASELINE configurations through the same code (one bench line each) -> gpurun_out/r02_bench_<cfg>_n1.json
for c in cfg1 cfg3 cfg4 cfg5; do
  python bench.py --workload $c --no-gpu-eager --no-cpu-baseline > gpurun_out/r02_bench_${c}_n1.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${c}_n1.json")); print("$c", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["train_step"]["ms_per_step"])
PY
done
python bench.py --workload cfg5 --mode infer --no-gpu-eager --no-cpu-baseline > gpurun_out/r02_bench_cfg5_infer_n1.json 2> gpurun_out/bench_cfg5i.err
python -c "import json; d=json.load(open('gpurun_out/r02_bench_cfg5_infer_n1.json')); print('cfg5 infer', d['ms_per_step'], d['value'])"
