# warp-specialised step kernel: A/B timing against the one-thread-per-env kernel
set -x
mkdir -p gpurun_out
rm -f gpurun_out/split_ab.log
V=$PWD/hlynr_intercept_b200/_variants
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -k "split_kernel" > gpurun_out/pytest_split.log 2>&1; tail -3 gpurun_out/pytest_split.log
HLYNR_OPTS=split=0 timeout 300 python tools/aged_time.py cfg4 fp32 2>&1 | tail -1 | tee -a gpurun_out/split_ab.log
for lib in "" $V/libhlynr_b200_ws80.so $V/libhlynr_b200_ws112.so; do
  HLYNR_B200_LIB=$lib HLYNR_OPTS=split=1 timeout 300 python tools/aged_time.py cfg4 fp32 2>&1 | tail -1 | tee -a gpurun_out/split_ab.log
done
for c in 4 5 6; do HLYNR_OPTS=split=1,ws_ctas_per_sm=$c timeout 300 python tools/aged_time.py cfg4 fp32 2>&1 | tail -1 | tee -a gpurun_out/split_ab.log; done
HLYNR_OPTS=split=1 timeout 300 python tools/aged_time.py cfg4 fp64 2>&1 | tail -1 | tee -a gpurun_out/split_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_ws80.so HLYNR_OPTS=split=1 timeout 300 python tools/aged_time.py cfg4 fp64 2>&1 | tail -1 | tee -a gpurun_out/split_ab.log
HLYNR_OPTS=split=1 ncu --set full --clock-control none --import-source on -k regex:step_kernel_ws -s 5 -c 1 -f -o gpurun_out/prof_ws python tools/aged_step.py cfg4 > gpurun_out/ncu_ws.log 2>&1
