#!/bin/bash
# round 2, GPU run 1: new parity tests + baseline instrumentation (device timeline, solver statistics, L2 residency) + configs 3/4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
nproc > gpurun_out/r1_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r1_host.txt 2>&1; python -c "import os;print(len(os.sched_getaffinity(0)))" >> gpurun_out/r1_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest.log
for spec in "Lift Panda 4096 OSC_POSE" "Stack Sawyer 8192 JOINT_VELOCITY" "NutAssemblyRound Panda 16384 OSC_POSE"; do
  B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 600 python tools/probe_instr.py $spec > gpurun_out/r1_instr_$(echo $spec | cut -d' ' -f1).log 2>&1
done
for g in 1 2; do B2S_GROUPS=$g B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r1_instr_Lift_G$g.log 2>&1; cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/instr_Lift_Panda_4096_G$g.json; done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_c2.json 2> gpurun_out/r1_bench_c2.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 3 --no-timeline > gpurun_out/r1_bench_c3.json 2> gpurun_out/r1_bench_c3.err
timeout 900 python bench.py --steps 6 --warmup 3 --config 4 --no-timeline > gpurun_out/r1_bench_c4.json 2> gpurun_out/r1_bench_c4.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 5 --no-timeline --no-cpu-baseline > gpurun_out/r1_bench_c5.json 2> gpurun_out/r1_bench_c5.err
# L2 residency of the workspace rows: same kernels with and without ncu's cache flush, lts hit rates + dram bytes
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum -k regex:"phase_kernel|narrow" -s 400 -c 16 --csv --log-file gpurun_out/r1_l2_nocachectl.csv python tools/probe_pipeline.py > gpurun_out/r1_l2a.log 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum -k regex:"phase_kernel|narrow" -s 400 -c 16 --csv --log-file gpurun_out/r1_l2_flush.csv python tools/probe_pipeline.py > gpurun_out/r1_l2b.log 2>&1
echo done
