# occupancy sensitivity of the fp32 step kernel (4 / 3 / 2 CTAs per SM through unused dynamic shared memory), fp64 fast-sqrt A/B, fp64 parity
set -x
mkdir -p gpurun_out
for pad in 0 50000 80000; do HLYNR_OCC_PAD_BYTES=$pad python tools/aged_time.py cfg4,cfg2 fp32 2>&1 | sed "s/^/pad=$pad /" | tee -a gpurun_out/occ_sweep.log; done
HLYNR_OCC_PAD_BYTES=50000 python tools/aged_time.py cfg2 fp32 4096 2>&1 | tee -a gpurun_out/occ_sweep.log
python tools/aged_time.py cfg4,cfg2,cfg3 fp64 2>&1 | tee gpurun_out/f64_sqrt_ab.log
HLYNR_B200_LIB=$PWD/hlynr_intercept_b200/_variants/libhlynr_b200_exactsqrt.so python tools/aged_time.py cfg4,cfg2,cfg3 fp64 2>&1 | tee -a gpurun_out/f64_sqrt_ab.log
timeout 1500 python -m pytest tests/test_cuda_parity.py -m gpu -q -x > gpurun_out/pytest_parity.log 2>&1; tail -5 gpurun_out/pytest_parity.log
