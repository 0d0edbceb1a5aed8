#!/bin/bash
# Round evidence on one B200 (run under gpurun from the repo root): tests, bench lines, ncu launch list and
# full captures of the two dominant kernels.  Outputs under gpurun_out/ev_*; numbers printed under ncu are not bench values.
set -u
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/ev_pytest.txt; cat $O/ev_pytest.txt
python bench.py > $O/ev_bench_default.json 2> $O/ev_bench_default.err; cut -c1-160 $O/ev_bench_default.json
python bench.py --streams 1024 --nn tensor --no-cpu-baseline > $O/ev_bench_1024_tensor.json 2>/dev/null
python bench.py --streams 1024 --nn fp32 --no-cpu-baseline > $O/ev_bench_1024_fp32.json 2>/dev/null
python bench.py --path traindata > $O/ev_bench_traindata.json 2>/dev/null; cut -c1-160 $O/ev_bench_traindata.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ev_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:analysis_kernel --launch-skip 3 -c 1 -f -o $O/ev_analysis \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ev_ncu_analysis.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel --launch-skip 134 -c 6 -f -o $O/ev_tc_gemm \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ev_ncu_tc.log 2>&1
ls -la $O/ev_*
