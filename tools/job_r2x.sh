# plane / ring-row pitch skew (L1 set conflicts between the planes of an env?)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/pad_skew_ab.log
for sk in 0 32 160 544 2080 0; do
HLYNR_PAD_SKEW=$sk timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed "s/^/skew=$sk: /" | tee -a gpurun_out/pad_skew_ab.log
done
