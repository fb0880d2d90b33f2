cd /root/repo; mkdir -p gpurun_out
for tw in 0 8 16 32 64 128; do
  echo "== TW=$tw"
  CVVAE_CONV_TW=$tw timeout 600 python tools/bench_conv.py --reps 3 2>&1 | grep '"layer"' | cut -c1-100
done 2>&1 | tee gpurun_out/tw_sweep_all.log
