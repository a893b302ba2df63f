# second compact layout (cfg3, domain randomization: drag peak in i0.y, no f3 plane): bit-identity test and A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -k "compact" 2>&1 | tail -2
rm -f gpurun_out/compact_dr_ab.log
for i in 1 2; do
HLYNR_NO_COMPACT=1 timeout 300 python tools/aged_time.py cfg3 fp32 2>&1 | tail -1 | sed 's/^/12 planes: /' | tee -a gpurun_out/compact_dr_ab.log
timeout 300 python tools/aged_time.py cfg3 fp32 2>&1 | tail -1 | sed 's/^/11 planes: /' | tee -a gpurun_out/compact_dr_ab.log
done
HLYNR_NO_COMPACT=1 timeout 300 python tools/aged_time.py cfg3 fp32 262144 2>&1 | tail -1 | sed 's/^/12 planes: /' | tee -a gpurun_out/compact_dr_ab.log
timeout 300 python tools/aged_time.py cfg3 fp32 262144 2>&1 | tail -1 | sed 's/^/11 planes: /' | tee -a gpurun_out/compact_dr_ab.log
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_vec_env.py tests/test_cuda_stats.py -m gpu -q -x -k "cfg3 or radar or stat or sharding or windows" 2>&1 | tail -2
