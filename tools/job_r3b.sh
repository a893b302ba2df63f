# ring prefetch addresses ready-made from the host (main) against computed in the kernel prologue (variant)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/prefetch_addr_ab.log
for i in 1 2; do
HLYNR_B200_LIB=$V/libhlynr_b200_pfold.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/computed in the kernel: /" | tee -a gpurun_out/prefetch_addr_ab.log
timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp32 2>&1 | tail -3 | sed "s/^/from the host: /" | tee -a gpurun_out/prefetch_addr_ab.log
done
timeout 300 python tools/aged_time.py cfg4,cfg3 fp64 2>&1 | tail -2 | sed "s/^/from the host: /" | tee -a gpurun_out/prefetch_addr_ab.log
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_vec_env.py -m gpu -q -x -k "golden or compact or sharding or graph or fused or pipelined" 2>&1 | tail -3
