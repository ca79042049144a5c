# A/B timing of kernel-variant builds on the GPU box:
#   make -C cusmc_b200/csrc VARIANT=x EXTRA=-DSOME_MACRO=1        (builds cusmc_b200/libcusmc_b200_x.so)
#   gpurun -- 'VARIANTS="_x" WHICH="c4 1000" bash profiles/run_ab.sh'
# "" is the default build.  Results: gpurun_out/pfb_$TAG.log
set -x; mkdir -p gpurun_out
TAG=${TAG:-ab}
for v in "" ${VARIANTS:-}; do
  for i in 1 2; do echo "variant '$v'"; CUSMC_B200_LIB=$PWD/cusmc_b200/libcusmc_b200$v.so python profiles/prof_c5.py ${WHICH:-c5 41}; done
done > gpurun_out/pfb_$TAG.log 2>&1
cat gpurun_out/pfb_$TAG.log
