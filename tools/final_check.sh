#!/bin/bash
# Round-end check on a GPU box: GPU tests, smoke, the default bench line, the ncu launch list of the eager bench command.
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_cfg2_n1_final.json 2>gpurun_out/bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_cfg2_n1_final.json")); print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["step_tensor_frac"], d["train_step"]["ms_per_step"], d["gpu_eager_baseline"]["ms_per_step"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
PY
CMD="python bench.py --graphs 0 --steps 1 --warmup 3 --no-train-step --no-gpu-eager --no-cpu-baseline --no-e2e --no-roofline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1800 -c 640 --csv --log-file gpurun_out/r02_launches_cfg2_final3.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
python tools/gemm_table.py > gpurun_out/r02_gemm_table_final.txt 2>/dev/null; head -3 gpurun_out/r02_gemm_table_final.txt
