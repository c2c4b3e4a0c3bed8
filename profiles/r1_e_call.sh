mkdir -p gpurun_out
for mb in 5 6; do SB200_SR_MINB=$mb python profiles/tune_ops.py --graph poisson --size 4096 --ops permute2d_rcm,permute2d_rand --reps 10 2>&1 | tail -1; done
