# delayed onboard ring load issued before the wind update (1), before the outcome (2), after the outcome (3), or inside observe() (main)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/oring_hoist_ab.log
for i in 1 2; do
for v in "" oring1 oring2 oring3; do
  if [ -z "$v" ]; then lib=""; else lib=$V/libhlynr_b200_$v.so; fi
  HLYNR_B200_LIB=$lib timeout 300 python tools/aged_time.py cfg4 fp32 2>&1 | tail -1 | sed "s/^/[${v:-main}] /" | tee -a gpurun_out/oring_hoist_ab.log
done
done
