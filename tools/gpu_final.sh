#!/bin/bash
# Last GPU visit of the round: full parity suite on the default build, bench with both K1 versions, launch list, full ncu capture.
tag=${1:-f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/${tag}_pytest.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
B200_MC=1 timeout 200 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_mc1.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_mc1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -c 20 -o gpurun_out/${tag}_bpic python tools/run_pictures.py --only 4 --reps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
B200_TRACE=gpurun_out/${tag}_trace.csv timeout 200 python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_trace_bench.json 2>> gpurun_out/${tag}_bench.err
python tools/timeline.py gpurun_out/${tag}_trace.csv --from 64 --to 224 | tee gpurun_out/${tag}_timeline.txt
timeout 200 python bench.py --workload c2_1080p_main_ra --no-cpu-baseline > gpurun_out/${tag}_bench_1080p.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_1080p.json
