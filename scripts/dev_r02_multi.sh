#!/bin/bash
# dev: multi-GPU pass on ONE box with $1 GPUs: band-mode parity at every rank count, then the config 5 / config 3 / config 2 lines.
# usage: dev_r02_multi.sh <n_gpus_on_box> "<rank counts, e.g. 2 4 8>"
NG=$1; COUNTS=${2:-$1}
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() { # n, label, args...
  n=$1; label=$2; shift 2
  if [ "$n" = "1" ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/r02_${label}_n${n}.json 2> gpurun_out/r02_${label}_n${n}.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/r02_${label}_n${n}.json 2> gpurun_out/r02_${label}_n${n}.err; fi
  echo "$label N=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_${label}_n${n}.json').read().strip().splitlines()[-1])
    print('   value %.1f %s, %.3f ms/step, e2e %.1f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value']))
except Exception as e:
    print('   no line:', e); print(open('gpurun_out/r02_${label}_n${n}.err').read()[-1500:])
PY
}
for n in $COUNTS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29400 + n)) tests/run_tiled_ranks.py > gpurun_out/r02_tiled_ranks_n${n}.log 2>&1
  echo "tiled parity N=$n rc=$?"; grep -E "TILED-RANKS-OK|Error|assert" gpurun_out/r02_tiled_ranks_n${n}.log | head -5
done
run 1 config5 --workload config5 --steps 3 --warmup 3
for n in $COUNTS; do run $n config5 --workload config5 --steps 3 --warmup 3; done
for n in $COUNTS; do :; done
run $n config3 --workload config3 --steps 4 --warmup 2
run $n config2 --steps 20 --warmup 5 --no-cpu-baseline
