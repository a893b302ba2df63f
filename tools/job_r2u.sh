set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_bench_contract.py -m gpu -q -k "compact or gpu_arm" 2>&1 | tail -3
rm -f gpurun_out/f64_ctas_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp64 2>&1 | tail -3 | sed 's/^/3 CTAs per SM: /' | tee -a gpurun_out/f64_ctas_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_f64b4.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp64 2>&1 | tail -3 | sed 's/^/4 CTAs per SM: /' | tee -a gpurun_out/f64_ctas_ab.log
done
