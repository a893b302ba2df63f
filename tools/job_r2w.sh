# delayed onboard ring sample loaded right after the interceptor update (variant) against inside observe() (main)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/early_oring_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed 's/^/in observe: /' | tee -a gpurun_out/early_oring_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_earlyo.so timeout 300 python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | tail -2 | sed 's/^/early: /' | tee -a gpurun_out/early_oring_ab.log
done
HLYNR_B200_LIB=$V/libhlynr_b200_earlyo.so timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "golden and cfg4 or compact or oracle_at" 2>&1 | tail -3
