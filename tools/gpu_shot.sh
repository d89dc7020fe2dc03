mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_comodgan_gpu.py tests/test_u8_gpu.py -q -s 2>&1 | tail -40 > gpurun_out/new_tests.log; tail -15 gpurun_out/new_tests.log
timeout 45 python bench.py --workload comodgan --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_comodgan.json 2> gpurun_out/bench_comodgan.err; tail -c 1500 gpurun_out/bench_comodgan.json; tail -3 gpurun_out/bench_comodgan.err
timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('MIGAN', d['value'], d['e2e'], d.get('e2e_u8'))"
timeout 60 python -m pytest tests/test_ops_gpu.py -q 2>&1 | tail -3
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
