# on top of the counters-first load order: quaternion plane earlier, Philox draws hoisted (1 = evasion, 2 = evasion + wind)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/load_order_ab.log
for i in 1 2; do
for v in "" quat d1 d2 qd2; do
  if [ -z "$v" ]; then lib=""; else lib=$V/libhlynr_b200_$v.so; fi
  HLYNR_B200_LIB=$lib timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/[${v:-main}] /" | tee -a gpurun_out/load_order_ab.log
done
done
timeout 300 python tools/aged_time.py cfg4,cfg2 fp64 2>&1 | tail -2 | sed "s/^/[main] /" | tee -a gpurun_out/load_order_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_d2.so timeout 300 python tools/aged_time.py cfg4,cfg2 fp64 2>&1 | tail -2 | sed "s/^/[d2] /" | tee -a gpurun_out/load_order_ab.log
