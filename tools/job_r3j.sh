# Philox rounds unrolled by 2 instead of fully (code size) against fully unrolled (main)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/philox_rolled_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed "s/^/[unrolled] /" | tee -a gpurun_out/philox_rolled_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_prolled.so timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed "s/^/[by 2] /" | tee -a gpurun_out/philox_rolled_ab.log
done
timeout 300 python tools/aged_time.py cfg4,cfg2 fp64 2>&1 | tail -2 | sed "s/^/[unrolled] /" | tee -a gpurun_out/philox_rolled_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_prolled.so timeout 300 python tools/aged_time.py cfg4,cfg2 fp64 2>&1 | tail -2 | sed "s/^/[by 2] /" | tee -a gpurun_out/philox_rolled_ab.log
