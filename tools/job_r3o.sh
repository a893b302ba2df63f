# delayed ground ring sample loaded before the ground detection / noise draws (variant) against after them (main; also: compact flag compile-time where impossible)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/gring_early_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/[main] /" | tee -a gpurun_out/gring_early_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_gearly.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/[early] /" | tee -a gpurun_out/gring_early_ab.log
done
HLYNR_B200_LIB=$V/libhlynr_b200_gearly.so timeout 300 python tools/aged_time.py cfg4 fp64 2>&1 | tail -1 | sed "s/^/[early] /" | tee -a gpurun_out/gring_early_ab.log
