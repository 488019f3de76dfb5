# Development aid: one bench.py run at N ranks with hard time limits; prints the key numbers of the JSON line.
N=${1:-8}; shift
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --watchdog-s 360 "$@" 2>gpurun_out/r02_n$N.err | grep '^{' > gpurun_out/r02_n$N.json
echo "exit ${PIPESTATUS[0]}"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_n$N.json').read().strip().split('\n')[-1])
print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['ddp'][:70], d['config']['cuda_graph'])
for k in ('rowshard','detmap','maptrv2_decoder'):
    if k in d: print(k, json.dumps(d[k])[:1500])
PY
grep -v "OMP_NUM\|\*\*\*\|^$\|UserWarning\|run_backward" gpurun_out/r02_n$N.err | tail -12 | cut -c1-220
