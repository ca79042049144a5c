# scratch driver for A/B timing runs on the GPU box (gpurun -- bash profiles/run_ab.sh)
set -x; mkdir -p gpurun_out
TAG=${TAG:-s4c}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
for i in 1 2 3; do python profiles/pf_breakdown.py c4 1000; done > gpurun_out/pfb_$TAG.log 2>&1
for i in 1 2 3; do CUSMC_B200_LIB=$PWD/cusmc_b200/libcusmc_b200_nopregen.so python profiles/pf_breakdown.py c4 1000; done >> gpurun_out/pfb_$TAG.log 2>&1
python profiles/pf_breakdown.py c5 41 >> gpurun_out/pfb_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pfb_$TAG.log
