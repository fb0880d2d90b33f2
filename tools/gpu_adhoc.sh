cd /root/repo; mkdir -p gpurun_out
for rep in 1 2; do
for na in 0 2; do
  echo "== NA=$na (rep $rep)"
  for l in "D 128->128 333 @17" "E 128->128 133" "D 256->128 333" "E 128->128 333 @17"; do
    CVVAE_CONV_NA=$na timeout 300 python tools/bench_conv.py --reps 5 --only "$l" 2>&1 | grep '"layer"' | cut -c1-100
  done
done
done 2>&1 | tee gpurun_out/na_sweep.log
