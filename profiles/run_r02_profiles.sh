# round-2 evidence pass on the GPU box (one gpurun call, ncu once):
#   gpurun --timeout 1500 -- 'bash profiles/run_r02_profiles.sh'
# 1. plain run of the profiled driver (must exit 0), 2. full-set capture of every hot kernel of the
# final build, 3. launch list of the bench command (cold-cache, serialised: compare SHARES).
set -x; mkdir -p gpurun_out
python profiles/prof_driver.py all > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"density_soa|pf_fused|tile_update|pf_persistent|mh_chains|mh_general|perpoint|metropolis|multinomial" -c 56 \
    -o gpurun_out/r02_final python profiles/prof_driver.py all > gpurun_out/r02_prof_ncu.log 2>&1
# the report of 56 full-set launches with sources is ~100 MB, above what travels back: export the pages the
# summaries are made from (profiles/ncu_summary.py, ncu_phases.py) and drop the report itself
ncu -i gpurun_out/r02_final.ncu-rep --page raw --csv > gpurun_out/r02_final_raw.csv 2> /dev/null
ncu -i gpurun_out/r02_final.ncu-rep --page source --csv --kernel-name 'regex:pf_fused_kernel<8, 1, 1, 0, 1, 1, 0>' \
    > gpurun_out/r02_final_source_pf_fused_d8.csv 2> /dev/null
ncu -i gpurun_out/r02_final.ncu-rep --page source --csv --kernel-name 'regex:pf_persistent' \
    > gpurun_out/r02_final_source_pf_persistent.csv 2> /dev/null
rm -f gpurun_out/r02_final.ncu-rep
python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_launches_ncu.log 2>&1
tail -n 2 gpurun_out/r02_prof_plain.log; tail -n 2 gpurun_out/r02_prof_ncu.log; ls -la gpurun_out/
