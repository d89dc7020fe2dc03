#!/bin/bash
# N-GPU shot (gpurun --gpus N -- bash tools/gpu_shot_multi.sh N [variants...]): bit-exact check of the sharded forward and
# bench lines.  Variant = ce | ce0 (no SMs left free) | ncclgather | nogather | 256.  Outputs in gpurun_out/.
N=${1:-2}; shift
VARIANTS=${@:-"ce ncclgather"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29530
FILES=""
if [ -z "$SKIP_BASE" ]; then   # SKIP_BASE=1: only the N-GPU bench lines (an 8-GPU box is charged 8x)
  timeout 200 $TR --master-port $PORT tools/multi_gpu_check.py ce 2>&1 | grep -E "multi_gpu_check|Error|error" | head -4
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s1.json 2> gpurun_out/s1.err
  FILES="s1"
fi
for v in $VARIANTS; do
  PORT=$((PORT+2))
  ENVV="X=1"
  case $v in
    ce)         ARGS="";;
    ce0)        ARGS=""; ENVV="MIGAN_TC_RESERVE_SMS=0";;
    ncclgather) ARGS="--gather nccl";;
    nogather)   ARGS="--no-gather";;
    256)        ARGS="--res 256";;
  esac
  env $ENVV timeout 200 $TR --master-port $((PORT+1)) bench.py --gpus $N $ARGS --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s${N}_$v.json 2> gpurun_out/s${N}_$v.err
  FILES="$FILES s${N}_$v"
done
for f in $FILES; do
  python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print("%-16s value %6.0f img/s  %7.3f ms/step  e2e %6.0f img/s | %s" % (f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["workload"][-90:]))
except Exception as e:
    print(f, "failed", e)
    print(open("gpurun_out/%s.err" % f).read()[-1200:])
PY
done
true
