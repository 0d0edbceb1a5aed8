#!/bin/bash
# Round evidence on one B200 (run under gpurun from the repo root): bench lines, ncu launch list and full captures of
# the dominant kernels.  Outputs under gpurun_out/ev2_*; numbers printed under ncu are not bench values.
set -u
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/ev2_bench_default.json 2> $O/ev2_bench_default.err; cut -c1-200 $O/ev2_bench_default.json
PNB_OVERLAP=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-int16-run > $O/ev2_bench_serial.json 2>/dev/null
python bench.py --frames 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/ev2_bench_f8.json 2>/dev/null
python bench.py --path xcorr --streams 65536 --steps 20 --warmup 5 > $O/ev2_bench_xcorr.json 2> $O/ev2_xcorr.err; cut -c1-200 $O/ev2_bench_xcorr.json
python bench.py --streams 1024 --frames 32 --nn tensor --no-cpu-baseline > $O/ev2_bench_1024_tensor.json 2>/dev/null
python bench.py --streams 1024 --frames 32 --nn fp32 --no-cpu-baseline > $O/ev2_bench_1024_fp32.json 2>/dev/null
python bench.py --streams 1024 --frames 8 --nn fp32 --no-cpu-baseline --no-int16-run > $O/ev2_bench_1024_fp32_f8.json 2>/dev/null
python bench.py --streams 16384 --frames 8 --nn fp32 --no-cpu-baseline --no-int16-run > $O/ev2_bench_16384_fp32.json 2>/dev/null
python bench.py --path traindata > $O/ev2_bench_traindata.json 2>/dev/null
timeout 600 ncu --kernel-name-base mangled -k regex:pnb --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/ev2_launches.csv \
  python bench.py --frames 8 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-int16-run > $O/ev2_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_chain_kernel --launch-skip 3 -c 1 -f -o $O/ev2_chain \
  python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-int16-run > $O/ev2_ncu_chain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:analysis_kernel --launch-skip 3 -c 1 -f -o $O/ev2_analysis \
  python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-int16-run > $O/ev2_ncu_analysis.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 6 -c 4 -f -o $O/ev2_dense \
  python bench.py --frames 8 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-int16-run > $O/ev2_ncu_dense.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pitch_only_kernel --launch-skip 3 -c 1 -f -o $O/ev2_pitch \
  python bench.py --path xcorr --streams 65536 --steps 1 --warmup 3 --no-cpu-baseline > $O/ev2_ncu_pitch.log 2>&1
timeout 600 ncu --kernel-name-base mangled -k regex:pnb --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ev2_launches_1024_fp32.csv \
  python bench.py --streams 1024 --frames 8 --nn fp32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-int16-run > $O/ev2_launches_1024_fp32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_chain_f32_kernel --launch-skip 3 -c 1 -f -o $O/ev2_chain_f32 \
  python bench.py --streams 1024 --frames 8 --nn fp32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-int16-run > $O/ev2_ncu_chain_f32.log 2>&1
python tools/latency_single_stream.py > $O/ev2_latency.json 2>/dev/null
python tools/timeline.py > $O/ev2_timeline.json 2>/dev/null
ls -la $O/ev2_*
