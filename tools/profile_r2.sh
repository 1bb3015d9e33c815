# Round-2 profiling commands (one gpurun call).  Each ncu run follows the same command exiting 0 without ncu.
# Summaries: python tools/ncu_summary.py {rep,list} ... -> profiles/r2_ncu_iteration.md, profiles/r2_launch_summary.csv
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c3 --no-canonical-spmv"
# 1. launch list of the default bench command (per-launch times are cold-cache and serialised: read SHARES)
$CMD > gpurun_out/r2_prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
# 2. --set full capture of every solver kernel once, on the resident 9.7 M-DoF system (cuProfilerStart/Stop bracket)
python tools/profile_kernels.py > gpurun_out/r2_pk_plain.log 2>&1 || { echo "plain kernel run failed"; exit 1; }
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_prof_iteration \
    python tools/profile_kernels.py > gpurun_out/r2_ncu_pk.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
