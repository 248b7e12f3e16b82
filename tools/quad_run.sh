mkdir -p gpurun_out/full1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/full1/pytest_gpu.log 2>&1
tail -5 gpurun_out/full1/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full1/smoke.log 2>&1; tail -3 gpurun_out/full1/smoke.log
timeout 200 python bench.py > gpurun_out/full1/cfg4_R.json 2> gpurun_out/full1/cfg4_R.err
timeout 200 python bench.py --objective marginal --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/full1/cfg4_M.json 2> gpurun_out/full1/cfg4_M.err
timeout 200 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/full1/cfg5_R.json 2> gpurun_out/full1/cfg5_R.err
timeout 300 python bench.py --workload cfg5 --objective marginal --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/full1/cfg5_M.json 2> gpurun_out/full1/cfg5_M.err
timeout 200 python bench.py --workload cfg3 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/full1/cfg3_R.json 2> gpurun_out/full1/cfg3_R.err
