# Round-2 profiling commands (one gpurun call): slab-512 A/B timing, ncu launch list and --set full captures of the
# assembly / Schur / Gram-Schmidt / slab kernels on the default bench command.  Each ncu run follows the same
# command exiting 0 without ncu.
P=navierstokes-capoferri_cecchettini_untila_b200
NSB_TIME_ONLY=sweep_F,block_spmv,g_apply NSB_LIBNSB=$PWD/$P/libnsb_t512.so python tools/time_kernels.py > gpurun_out/r2_slab_t512.log 2>&1; tail -1 gpurun_out/r2_slab_t512.log
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c3 --no-canonical-spmv"
$CMD > gpurun_out/r2_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 900 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"assemble_cells_kernel|schur_outer_kernel" -c 2 -o gpurun_out/r2_prof_asm_schur $CMD > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ortho_kernel -s 45 -c 3 -o gpurun_out/r2_prof_ortho $CMD > gpurun_out/r2_ncu_o.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fs_slab_sweep_kernel|fs_slab_apply_kernel" -s 30 -c 2 -o gpurun_out/r2_prof_sweep $CMD > gpurun_out/r2_ncu_s.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
