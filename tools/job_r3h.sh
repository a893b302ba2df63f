set -x
mkdir -p gpurun_out; rm -f gpurun_out/prefetch_waves.log
for w in 0 1 2 3; do HLYNR_OPTS=prefetch_waves=$w timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | tee -a gpurun_out/prefetch_waves.log; done
