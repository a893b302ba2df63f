# delayed ring rows: no prefetch / prefetch.global.L1 (main) / prefetch.global.L2
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/ring_prefetch_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/L1: /" | tee -a gpurun_out/ring_prefetch_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_ringpf0.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/none: /" | tee -a gpurun_out/ring_prefetch_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_ringpf2.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/L2: /" | tee -a gpurun_out/ring_prefetch_ab.log
done
