#!/bin/bash
# One GPU visit: parity suite, bench (4K + 1080p), the real decoder side by side, launch list.  Run through gpurun:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
tag=${1:-r}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ${PYTEST_K:+-k "$PYTEST_K"} ) > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
timeout 300 python bench.py --workload c2_1080p_main_ra --no-cpu-baseline > gpurun_out/${tag}_bench_1080p.json 2>> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench_1080p.json
if [ -z "$SKIP_DECODE" ]; then
  timeout 900 python tools/decode_bench.py --repeat 1 --passes 3 > gpurun_out/${tag}_decode.json 2> gpurun_out/${tag}_decode.err
  cat gpurun_out/${tag}_decode.json
  if [ -f oracle/_ref/streams/c3_4k_calm_17.hevc ]; then    # lightly coded 4K stream (tools/make_bench_streams.sh): the bit rate of real content
    timeout 600 python tools/decode_bench.py oracle/_ref/streams/c3_4k_calm_17.hevc --repeat 1 --passes 4 > gpurun_out/${tag}_decode_calm.json 2>> gpurun_out/${tag}_decode.err
    cat gpurun_out/${tag}_decode_calm.json
  fi
fi
if [ -n "$WITH_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
fi
if [ -n "$WITH_NCU_FULL" ]; then   # one B picture (blob 4), second repetition: every kernel of the picture with the full set
  timeout 600 ncu --set full --clock-control none --import-source on -c 26 -o gpurun_out/${tag}_bpic python tools/run_pictures.py --only 4 --reps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
fi
if [ -n "$WITH_PCIE" ]; then B200_VERBOSE=1 python tools/pcie_probe.py > gpurun_out/${tag}_pcie.txt 2>&1; cat gpurun_out/${tag}_pcie.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/${tag}_pcie.txt 2>&1; nproc >> gpurun_out/${tag}_pcie.txt; fi
if [ -n "$WITH_TRACE" ]; then
  B200_TRACE=gpurun_out/${tag}_trace.csv timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline > gpurun_out/${tag}_trace_bench.json 2>> gpurun_out/${tag}_bench.err
  python tools/timeline.py gpurun_out/${tag}_trace.csv --from 64 --to 224 | tee gpurun_out/${tag}_timeline.txt
fi
if [ -n "$WITH_VARIANTS" ]; then   # the switches prepared without a GPU: parity subset + bench line each
  for v in "B200_MC=3" "B200_MC=3 B200_MC_DESC=1" "B200_MC_DESC=1" "B200_EDGES_SPARSE=1" "B200_MC=3 B200_MC_DESC=1 B200_EDGES_SPARSE=1"; do
    name=$(echo "$v" | tr ' =' '__')
    ( env $v timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2 ) > gpurun_out/${tag}_var_${name}_pytest.log
    env $v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${tag}_var_${name}_bench.json 2>> gpurun_out/${tag}_bench.err
    echo "$v: $(cat gpurun_out/${tag}_var_${name}_pytest.log | tail -1) $(cat gpurun_out/${tag}_var_${name}_bench.json | cut -c1-200)"
  done
fi
