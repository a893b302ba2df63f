# counters' planes loaded first / Philox draws hoisted ahead of the interceptor update
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/draws_early_ab.log
for i in 1 2; do
for v in "" ctrfirst drawe1 drawe2 ctrdraw; do
  if [ -z "$v" ]; then lib=""; else lib=$V/libhlynr_b200_$v.so; fi
  HLYNR_B200_LIB=$lib timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed "s/^/[${v:-main}] /" | tee -a gpurun_out/draws_early_ab.log
done
done
