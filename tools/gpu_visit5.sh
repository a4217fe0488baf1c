#!/bin/bash
# 2-GPU visit: multi-GPU parity, scaling bench at N=1 and N=2, picture timeline per rank
tag=${1:-b5}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/${tag}_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/verify_multi_gpu.py --gops 8 2>&1 | tail -3 | tee gpurun_out/${tag}_verify2.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "frame_parallel_over_two" 2>&1 | tail -3 | tee -a gpurun_out/${tag}_verify2.txt
timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-stream > gpurun_out/${tag}_bench_n1.json 2>> gpurun_out/${tag}_bench.err
cut -c1-300 gpurun_out/${tag}_bench_n1.json
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --no-stream > gpurun_out/${tag}_bench_n2.json 2>> gpurun_out/${tag}_bench.err
cut -c1-300 gpurun_out/${tag}_bench_n2.json
B200_TRACE=gpurun_out/${tag}_trace_gpu%d.csv timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 24 --warmup 4 --no-cpu-baseline --no-stream > gpurun_out/${tag}_bench_n2_trace.json 2>> gpurun_out/${tag}_bench.err
for d in 0 1; do python tools/timeline.py gpurun_out/${tag}_trace_gpu$d.csv --from 64 --to 224 > gpurun_out/${tag}_timeline_gpu$d.txt 2>&1; tail -12 gpurun_out/${tag}_timeline_gpu$d.txt; done
# the real decoder as N replicas (one per GPU, host threads split)
timeout 900 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 64 --no-cpu-baseline > gpurun_out/${tag}_bench_n2_stream.json 2>> gpurun_out/${tag}_bench.err
python -c "import json;l=open(\"gpurun_out/${tag}_bench_n2_stream.json\").read().strip().splitlines();print(\"n2 e2e\", json.loads(l[-1])[\"e2e\"] if l else None)"
