# own-motion channels emitted between the delayed ring loads and their first use (variant) against at the end of observe() (main)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/own_channels_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/at the end: /" | tee -a gpurun_out/own_channels_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_ownearly.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/early: /" | tee -a gpurun_out/own_channels_ab.log
done
HLYNR_B200_LIB=$V/libhlynr_b200_ownearly.so timeout 300 python tools/aged_time.py cfg4 fp64 2>&1 | tail -1 | sed "s/^/early: /" | tee -a gpurun_out/own_channels_ab.log
