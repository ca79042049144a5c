# end-of-session pass on the GPU box: tests, bench lines (ours + reference arm), ncu launch list
# usage: gpurun --timeout 900 -- 'TAG=r01s bash profiles/run_final.sh'
set -x; mkdir -p gpurun_out
TAG=${TAG:-r01s}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --quick > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.json
