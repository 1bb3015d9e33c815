# Round-2 profiling commands (one gpurun call): ncu launch list and --set full captures of the assembly / Schur /
# Gram-Schmidt / slab / CSR kernels on the default bench command.  Each ncu run follows the same command exiting 0
# without ncu.  Summaries: python tools/ncu_summary.py (committed under profiles/).
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-c3 --no-canonical-spmv"
$CMD > gpurun_out/r2_prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1200 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"assemble_cells_kernel|schur_outer_kernel" -c 2 -o gpurun_out/r2_prof_asm_schur -f $CMD > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ortho_kernel<(16|20|24|28)" -s 4 -c 3 -o gpurun_out/r2_prof_ortho -f $CMD > gpurun_out/r2_ncu_o.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fs_slab_sweep_kernel|fs_slab_apply_kernel" -s 40 -c 2 -o gpurun_out/r2_prof_sweep -f $CMD > gpurun_out/r2_ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"g_slab_apply_kernel|spmv_kernel<32|cheb_sweep_kernel<8" -s 30 -c 4 -o gpurun_out/r2_prof_csr -f $CMD > gpurun_out/r2_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
