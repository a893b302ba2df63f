mkdir -p gpurun_out
for rep in 1 2; do
HLYNR_B200_LIB=$PWD/hlynr_intercept_b200/_exp_base.so python tools/aged_time.py cfg4,cfg2,cfg3 1 2>&1 | tee -a gpurun_out/cold_split.log
HLYNR_B200_LIB=$PWD/hlynr_intercept_b200/_exp_cold.so python tools/aged_time.py cfg4,cfg2,cfg3 1 2>&1 | tee -a gpurun_out/cold_split.log
done
HLYNR_B200_LIB=$PWD/hlynr_intercept_b200/_exp_cold.so python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "4096 or sharding or golden" 2>&1 | tail -3
