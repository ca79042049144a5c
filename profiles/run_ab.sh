# scratch driver for A/B timing runs on the GPU box (gpurun -- bash profiles/run_ab.sh)
set -x; mkdir -p gpurun_out
TAG=${TAG:-s4i}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
for i in 1 2; do python profiles/pf_breakdown.py c5 41; done > gpurun_out/pfb_$TAG.log 2>&1
for i in 1 2; do python profiles/pf_breakdown.py c4 1000; done >> gpurun_out/pfb_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pfb_$TAG.log
