# last evidence pass of round 2 (one gpurun call): full-set capture of the kernels that changed after
# run_r02_profiles.sh (density with the dependent-launch prologue, fused step incl. dense / Student-t, persistent
# kernel, Metropolis), raw page exported on the box (the report itself does not travel).
#   gpurun --timeout 900 -- 'bash profiles/run_r02i_profiles.sh'
set -x; mkdir -p gpurun_out
for part in logpdf pf pfmore; do
  python profiles/prof_driver.py $part > gpurun_out/r02i_plain_$part.log 2>&1 || exit 1
done
for part in logpdf pf pfmore; do
  ncu --set full --clock-control none --import-source on \
      -k regex:"density_soa|pf_fused|tile_update|pf_persistent|multinomial" -c 24 -f \
      -o gpurun_out/r02i_$part python profiles/prof_driver.py $part > gpurun_out/r02i_ncu_$part.log 2>&1
  ncu -i gpurun_out/r02i_$part.ncu-rep --page raw --csv > gpurun_out/r02i_raw_$part.csv 2> /dev/null
done
ncu -i gpurun_out/r02i_pf.ncu-rep --page source --csv --kernel-name 'regex:pf_persistent' > gpurun_out/r02i_source_pf_persistent.csv 2> /dev/null
rm -f gpurun_out/r02i_*.ncu-rep
ls -la gpurun_out/
