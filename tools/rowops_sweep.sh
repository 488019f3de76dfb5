#!/bin/bash
# One GPU call: correctness of a variant build of the row kernels (the dropout / module tests), then the
# micro-benchmark of every variant library given on the command line (the first one also over a sweep of
# row counts: fixed cost against per-row cost).   tools/rowops_sweep.sh w1 base v1 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
first=$1
APOLLO_B200_LIB=build/libmsda_$first.so timeout 600 python -m pytest tests/test_dropout_gpu.py tests/test_modules_gpu.py -m gpu -x -q > gpurun_out/rowops_tests_$first.log 2>&1
echo "tests($first) rc=$?"; tail -3 gpurun_out/rowops_tests_$first.log
: > gpurun_out/rowops_sweep.jsonl
APOLLO_B200_LIB=build/libmsda_$first.so timeout 300 python tools/rowops_bench.py --tag $first --rows 10000,20000,80000,160000 >> gpurun_out/rowops_sweep.jsonl 2>gpurun_out/rowops_bench_rows.err
for tag in "$@"; do
  APOLLO_B200_LIB=build/libmsda_$tag.so timeout 300 python tools/rowops_bench.py --tag $tag >> gpurun_out/rowops_sweep.jsonl 2>gpurun_out/rowops_bench_$tag.err || echo "bench $tag failed"
done
python - <<'PY'
import json
for l in open('gpurun_out/rowops_sweep.jsonl'):
    d = json.loads(l)
    print(d['tag'], d['rows'], ' '.join(f"{k.replace('_dropout','_dr').replace('residual','res')}={v['us']}" for k, v in d.items() if isinstance(v, dict)))
PY
