#!/bin/bash
tag=${1:-b6}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_stream_dropin.py -m gpu -x -q -k "baseline_shape or ctb_granular" --durations=5 ) > gpurun_out/${tag}_pytest_recipes.log 2>&1
tail -8 gpurun_out/${tag}_pytest_recipes.log
( B200_MC=4 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3 ) | tee gpurun_out/${tag}_pytest_mc4.log
for v in "B200_MC=1" "B200_MC=4 B200_MC4_CTAS=2" "B200_MC=4 B200_MC4_CTAS=3" "B200_MC=4 B200_MC4_CTAS=4" "B200_MC=4 B200_MC4_CTAS=6"; do
  name=$(echo "$v" | tr ' =' '__')
  env $v timeout 400 python bench.py --no-cpu-baseline --no-stream > gpurun_out/${tag}_var_${name}.json 2>> gpurun_out/${tag}_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_var_${name}.json").read().strip().splitlines()[-1])
print("$v", "value %.0f" % d["value"], {k: round(x["ms"]*1000,1) for k,x in d["roofline"]["stages"].items()})
PY
done
timeout 900 python tools/dbk_sao_sweep.py > gpurun_out/${tag}_sweep_c5.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_sweep_c5.json
timeout 600 python tools/dbk_sao_sweep.py --workload c3_4k_main10_ra > gpurun_out/${tag}_sweep_c3.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_sweep_c3.json
# ncu: launch list of the bench (default kernels), then every kernel of one B picture with the full set
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-stream > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -c 30 -o gpurun_out/${tag}_bpic python tools/run_pictures.py --only 4 --reps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out/${tag}_bpic.ncu-rep
