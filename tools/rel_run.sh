O=gpurun_out/rel; mkdir -p $O
GMVAE_CHAIN_ABL=8 timeout 300 python -m pytest tests/test_step_gpu.py -x -q -m gpu -k "multi_row or full_size or launch_plans or graph" > $O/pytest_rel.log 2>&1; tail -2 $O/pytest_rel.log
for rep in 1 2; do for a in 0 8; do
  GMVAE_CHAIN_ABL=$a python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/cfg4_a${a}_$rep.json 2>/dev/null
done; done
for a in 0 8; do
  GMVAE_CHAIN_ABL=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 131072 > $O/big_a$a.json 2>/dev/null
  GMVAE_CHAIN_ABL=$a python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg5 > $O/cfg5_a$a.json 2>/dev/null
done
