set -u
O=gpurun_out/ev3; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/cfg4_reference_arm.json 2> $O/cfg4_reference_arm.err
python bench.py > $O/cfg4_R.json 2> $O/cfg4_R.err
python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg5_R.json 2>/dev/null
python bench.py --batch 131072 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg4model_b131072.json 2>/dev/null
python bench.py --objective marginal --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg4_M.json 2>/dev/null
python tools/jobstat_chain.py 16384 cfg4 > $O/jobstat_cfg4.txt 2>&1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
  --log-file $O/launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_chain_kernel -s 4 -c 1 -f -o $O/prof_r2c_chain_cfg4 \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_full_cfg4.log 2>&1
