mkdir -p gpurun_out/abl4
for a in 0 1 4 5; do
  GMVAE_CHAIN_ABL=$a python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/abl4/cfg4_a$a.json 2> gpurun_out/abl4/cfg4_a$a.err
  GMVAE_CHAIN_ABL=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 131072 > gpurun_out/abl4/big_a$a.json 2> gpurun_out/abl4/big_a$a.err
  GMVAE_CHAIN_ABL=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/abl4/cfg5_a$a.json 2> gpurun_out/abl4/cfg5_a$a.err
done
GMVAE_VERBOSE=1 GMVAE_CHAIN_QUAD=1 GMVAE_CHAIN_ABL=5 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 131072 > gpurun_out/abl4/bigquad_a5.json 2> gpurun_out/abl4/bigquad_a5.err
python tools/jobstat_chain.py 16384 cfg4 > gpurun_out/abl4/jobstat_cfg4.txt 2>&1
python tools/jobstat_chain.py 131072 cfg4 > gpurun_out/abl4/jobstat_big.txt 2>&1
