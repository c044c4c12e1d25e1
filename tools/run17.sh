#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r17_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r17_pytest.log
tail -5 gpurun_out/r17_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r17_bench_c2.json 2> gpurun_out/r17_bench_c2.err; cut -c1-400 gpurun_out/r17_bench_c2.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r17_bench_ref.json 2> gpurun_out/r17_bench_ref.err; cut -c1-300 gpurun_out/r17_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/r17_launches.csv python bench.py --steps 1 --warmup 3 --preroll 60 --no-cpu-baseline --no-timeline > gpurun_out/r17_ncu_bench.log 2>&1
python tools/ncu_agg_launches.py gpurun_out/r17_launches.csv 2>&1 | cut -c1-200 | tail -12
