O=gpurun_out/ev4; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py > $O/cfg4_R.json 2> $O/cfg4_R.err; python tools/bench_summary.py "$O/*.json"
