# fp64 build: shared out-of-line Philox + normals / pow against the inlined copies; DRAM traffic of the compact cfg4 layout
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/f64_small_code_ab.log
for i in 1 2; do
timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp64 2>&1 | tail -3 | sed 's/^/shared copies: /' | tee -a gpurun_out/f64_small_code_ab.log
HLYNR_B200_LIB=$V/libhlynr_b200_f64inline.so timeout 300 python tools/aged_time.py cfg4,cfg2,cfg3 fp64 2>&1 | tail -3 | sed 's/^/inlined: /' | tee -a gpurun_out/f64_small_code_ab.log
done
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_kernel -s 1 -c 12 --csv --log-file gpurun_out/traffic_cfg4_compact.csv python tools/aged_step.py cfg4 > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu_traffic.log
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -q -x -k "oracle or windows" > gpurun_out/pytest_f64.log 2>&1; tail -3 gpurun_out/pytest_f64.log
