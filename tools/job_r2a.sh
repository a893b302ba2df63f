set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -40 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err ) 2>&1 | tail -3; tail -1 gpurun_out/bench.log | cut -c1-300; tail -5 gpurun_out/bench.err
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1 ) 2>&1 | tail -3; tail -1 gpurun_out/bench_ref.log | cut -c1-300
# DRAM traffic of 12 back-to-back steady-state launches, caches NOT flushed between them
python tools/aged_step.py cfg4 > gpurun_out/aged.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_kernel -s 1 -c 12 --csv --log-file gpurun_out/traffic_cfg4.csv python tools/aged_step.py cfg4 > gpurun_out/ncu_traffic.log 2>&1
tail -3 gpurun_out/ncu_traffic.log
for p in fp32 fp64; do python tools/aged_time.py cfg4,cfg2,cfg3 $p >> gpurun_out/aged_time.log 2>&1; done
for t in f64mb3 f64mb4; do HLYNR_B200_LIB=hlynr_intercept_b200/_variants/libhlynr_b200_$t.so python tools/aged_time.py cfg4,cfg2,cfg3 fp64 >> gpurun_out/aged_time.log 2>&1; done
cat gpurun_out/aged_time.log
ls -la gpurun_out
