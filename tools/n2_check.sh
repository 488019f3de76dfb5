# Development aid: bench.py at N = 2 with hard time limits (a hung collective costs GPU-minutes on every rank).
run() { timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --watchdog-s 120 "$@" 2>gpurun_out/r02_n2.err | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['ddp'][:60], d['config']['cuda_graph'])
for k in ('rowshard','detmap'):
    if k in d: print(k, json.dumps(d[k])[:1400])"; echo "exit $?"; grep -v "OMP_NUM\|\*\*\*\|^$" gpurun_out/r02_n2.err | tail -25 | cut -c1-200; }
for spec in "$@"; do echo "== $spec"; run $spec; done
