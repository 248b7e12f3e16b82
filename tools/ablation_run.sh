# Timing-only ablations of the chained kernel (profiles/r2_ablation_loads.md).  Build first with
#   GMVAE_NVCC_FLAGS=-DGMVAE_CHAIN_ABL python -m gmvae_b200.build --force
# then run this under gpurun; rebuild without the flag afterwards (`python -m gmvae_b200.build --force`).
# GMVAE_CHAIN_ABL bits: 1 = the epilogue warps skip their work, 4 = no tcgen05.mma is issued (results are garbage, the schedule is real).
O=gpurun_out/ablation; mkdir -p $O
for a in 0 1 4 5; do
  GMVAE_CHAIN_ABL=$a python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $O/cfg4_a$a.json 2> $O/cfg4_a$a.err
  GMVAE_CHAIN_ABL=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 131072 > $O/big_a$a.json 2> $O/big_a$a.err
  GMVAE_CHAIN_ABL=$a python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5 > $O/cfg5_a$a.json 2> $O/cfg5_a$a.err
done
# scheduling variants (production build is enough for these)
for v in "GMVAE_CHAIN_QUAD=1" "GMVAE_DEBUG_FLAGS=262144" "GMVAE_CHAIN_SPLIT=44" "GMVAE_CHAIN_BN=128"; do
  env $v python bench.py --steps 100 --warmup 10 --no-cpu-baseline > "$O/cfg4_$v.json" 2> /dev/null
done
python tools/bench_summary.py "$O/*.json"
