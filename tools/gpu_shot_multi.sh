#!/bin/bash
# N-GPU shot (gpurun --gpus N -- bash tools/gpu_shot_multi.sh N): bit-exact check of the sharded forward for both ready signals,
# one-GPU and N-GPU bench lines, the pipeline GPU tests.  Outputs land in gpurun_out/.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29511 tools/multi_gpu_check.py ce 2>&1 | grep -E "multi_gpu_check|Error|error" | head -5
MIGAN_CE_SIGNAL=nccl timeout 200 $TR --master-port 29512 tools/multi_gpu_check.py ce 2>&1 | grep -E "multi_gpu_check|Error|error" | head -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s1.json 2> gpurun_out/s1.err
timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s${N}_memops.json 2> gpurun_out/s${N}_memops.err
MIGAN_CE_SIGNAL=nccl timeout 200 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s${N}_nccl.json 2> gpurun_out/s${N}_nccl.err
if [ "$2" == "full" ]; then
  timeout 200 $TR --master-port 29515 bench.py --gpus $N --gather nccl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s${N}_ncclgather.json 2> gpurun_out/s${N}_ncclgather.err
  timeout 200 $TR --master-port 29516 bench.py --gpus $N --no-gather --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s${N}_nogather.json 2> gpurun_out/s${N}_nogather.err
  timeout 200 $TR --master-port 29517 bench.py --gpus $N --res 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s${N}_256.json 2> gpurun_out/s${N}_256.err
fi
for f in s1 s${N}_memops s${N}_nccl s${N}_ncclgather s${N}_nogather s${N}_256; do
  [ -f gpurun_out/$f.json ] && python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "value %.0f img/s" % d["value"], "%.3f ms/step" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "|", d["config"]["workload"][-70:])
except Exception as e:
    print(f, "failed", e)
    print(open("gpurun_out/%s.err" % f).read()[-1500:])
PY
done
[ "$3" == "tests" ] && timeout 300 python -m pytest tests/test_pipeline_gpu.py -x -q 2>&1 | tail -15
true
