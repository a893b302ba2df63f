mkdir -p gpurun_out
python tools/aged_time.py cfg4,cfg2,cfg3 1,3 2>&1 | tee gpurun_out/aged_time.log
HLYNR_B200_LIB=$PWD/hlynr_intercept_b200/_exp_b64.so python tools/aged_time.py cfg4,cfg2,cfg3 1 2>&1 | tee -a gpurun_out/aged_time.log
HLYNR_B200_LIB=$PWD/hlynr_intercept_b200/_exp_b32.so python tools/aged_time.py cfg4,cfg2,cfg3 1 2>&1 | tee -a gpurun_out/aged_time.log
