#!/bin/bash
# Round-end evidence run (one GPU): bench lines, ncu captures at the bench configuration (32 images), latency, Co-Mod-GAN.
# Everything lands in gpurun_out/; tools/summarize_profiles.py (run on the build box) turns it into profiles/.
#   gpurun --timeout 2400 -- tools/collect_profiles.sh
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r02_smi_before.csv
timeout 400 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r02_event_launches_migan512_bs32.json > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err
timeout 200 python bench.py --res 256 --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r02_event_launches_migan256_bs32.json > gpurun_out/r02_bench_migan256_bs32.json 2> gpurun_out/r02_bench_256.err
timeout 200 python bench.py --res 256 --path tc_fast --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_migan256_bs32_tc_fast.json 2>> gpurun_out/r02_bench_256.err
timeout 200 python tools/latency.py --res 512 --out gpurun_out/r02_latency_512.json > gpurun_out/r02_latency_512.log 2>&1
timeout 200 python tools/latency.py --res 256 --out gpurun_out/r02_latency_256.json > gpurun_out/r02_latency_256.log 2>&1
# ncu: every launch of one forward (cold-cache, serialised: compare SHARES) ...
ncu --metrics gpu__time_duration.sum --clock-control none -s 51 -c 51 --csv --log-file gpurun_out/r02_ncu_launch_list_migan512_bs32.csv python tools/ncu_target.py > /dev/null 2>&1
if [ -n "$QUICK" ]; then   # QUICK=1: bench lines, latency and the launch list only (the full captures take ~5 GPU-minutes)
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r02_smi_after.csv
  ls -la gpurun_out | tail -20
  exit 0
fi
# ... and full captures per kernel class, second forward, 32 images
timeout 500 $NCU -k regex:sepconv_tc -s 32 -c 5 -f -o gpurun_out/r02_ncu_tc_enc python tools/ncu_target.py > gpurun_out/ncu_a.log 2>&1
timeout 600 $NCU -k regex:sepconv_tc -s 56 -c 8 -f -o gpurun_out/r02_ncu_tc_syn python tools/ncu_target.py > gpurun_out/ncu_b.log 2>&1
timeout 400 $NCU -k regex:"dw3x3_down|torgb_img|dw3x3_act|up2_noise" -s 18 -c 18 -f -o gpurun_out/r02_ncu_ew python tools/ncu_target.py > gpurun_out/ncu_c.log 2>&1
for f in r02_ncu_tc_enc r02_ncu_tc_syn r02_ncu_ew; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/$f.source.csv 2>/dev/null
done
# the reports and source pages (~120 MB) do not fit gpurun's 64 MiB return channel: summarise the source view here, keep the raw pages
python tools/summarize_profiles.py --lines-only
rm -f gpurun_out/*.ncu-rep gpurun_out/*.source.csv
# Co-Mod-GAN: exact fp32 path (default) and the tcgen05 route
timeout 300 python bench.py --workload comodgan --steps 5 --warmup 3 > gpurun_out/r02_bench_comodgan256_bs16.json 2> gpurun_out/r02_bench_comod.err
timeout 200 python bench.py --workload comodgan --comod-gemm tc --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_comodgan256_bs16_tc.json 2>> gpurun_out/r02_bench_comod.err
timeout 300 python -m pytest tests/test_staged_gpu.py tests/test_comodgan_gpu.py -q -s 2>&1 | grep -E "max-abs|passed|failed" > gpurun_out/r02_comodgan_gpu_tests.log
cuobjdump -sass mi-gan_b200/lib/libmigan_b200.so | grep -oE "UTC[A-Z]*MMA[.A-Z0-9]*|UTMALDG[.0-9A-Z]*|UTMASTG[.0-9A-Z]*|UTMAPF[.A-Z0-9]*|LDTM[.x0-9]*|UTCBAR[.A-Z0-9]*|FFMA2|FMUL2|FADD2|FMNMX3|STG.E.ENL2.256|NANOSLEEP.SYNCS|SYNCS.PHASECHK.TRANS64.TRYWAIT" | sort | uniq -c | sort -rn > gpurun_out/r02_sass_mnemonics.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/r02_smi_after.csv
ls -la gpurun_out | tail -40
