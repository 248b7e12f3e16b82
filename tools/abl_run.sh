set -x
mkdir -p gpurun_out/abl3
for a in 0 1 2 4 3 5 6 7; do
  GMVAE_CHAIN_ABL=$a python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/abl3/cfg4_a$a.json 2> gpurun_out/abl3/cfg4_a$a.err
  GMVAE_CHAIN_ABL=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 131072 > gpurun_out/abl3/big_a$a.json 2> gpurun_out/abl3/big_a$a.err
  GMVAE_CHAIN_ABL=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/abl3/cfg5_a$a.json 2> gpurun_out/abl3/cfg5_a$a.err
done
