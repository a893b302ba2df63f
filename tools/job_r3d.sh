# lock-tick reduction ahead of the output stores (main) against after them (variant)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/tail_order_ab.log
for i in 1 2; do
HLYNR_B200_LIB=$V/libhlynr_b200_tailold.so timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed "s/^/after the stores: /" | tee -a gpurun_out/tail_order_ab.log
timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed "s/^/ahead of the stores: /" | tee -a gpurun_out/tail_order_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_tailold.so timeout 300 python tools/aged_time.py cfg4,cfg2 fp64 2>&1 | tail -2 | sed "s/^/after the stores: /" | tee -a gpurun_out/tail_order_ab.log
timeout 300 python tools/aged_time.py cfg4,cfg2 fp64 2>&1 | tail -2 | sed "s/^/ahead of the stores: /" | tee -a gpurun_out/tail_order_ab.log
done
