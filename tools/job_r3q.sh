# CTA size of the step kernel (128 main / 64 / 32 threads) at the BASELINE sizes, small batches in a CUDA graph
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/cta_size_ab.log
for v in "" blk64 blk32; do
  if [ -z "$v" ]; then lib=""; else lib=$V/libhlynr_b200_$v.so; fi
  HLYNR_B200_LIB=$lib timeout 300 python tools/baseline_sizes.py 2>&1 | tail -3 | sed "s/^/[${v:-128}] /" | cut -c1-260 | tee -a gpurun_out/cta_size_ab.log
  HLYNR_B200_LIB=$lib timeout 300 python tools/aged_time.py cfg4 fp32 131072 2>&1 | tail -1 | sed "s/^/[${v:-128}] /" | tee -a gpurun_out/cta_size_ab.log
done
