# prologue issue order: demand loads before the prefetches (variant) against prefetches first (main)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/prefetch_order_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed 's/^/prefetch first: /' | tee -a gpurun_out/prefetch_order_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_pflate.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed 's/^/loads first: /' | tee -a gpurun_out/prefetch_order_ab.log
done
timeout 300 python tools/aged_time.py cfg4 fp32 131072 2>&1 | tail -1 | sed 's/^/prefetch first: /' | tee -a gpurun_out/prefetch_order_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_pflate.so timeout 300 python tools/aged_time.py cfg4 fp32 131072 2>&1 | tail -1 | sed 's/^/loads first: /' | tee -a gpurun_out/prefetch_order_ab.log
timeout 300 python tools/aged_time.py cfg4,cfg3 fp64 2>&1 | tail -2 | sed 's/^/prefetch first: /' | tee -a gpurun_out/prefetch_order_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_pflate.so timeout 300 python tools/aged_time.py cfg4,cfg3 fp64 2>&1 | tail -2 | sed 's/^/loads first: /' | tee -a gpurun_out/prefetch_order_ab.log
