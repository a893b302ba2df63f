# compact plane layout (cfg4, fp32): A/B against the 11-plane layout, then the GPU suite
set -x
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl gpurun_out/compact_ab.log
for i in 1 2; do
HLYNR_NO_COMPACT=1 timeout 300 python tools/aged_time.py cfg4 fp32 2>&1 | tail -1 | sed 's/^/11 planes: /' | tee -a gpurun_out/compact_ab.log
timeout 300 python tools/aged_time.py cfg4 fp32 2>&1 | tail -1 | sed 's/^/10 planes: /' | tee -a gpurun_out/compact_ab.log
done
timeout 300 python tools/aged_time.py cfg4 fp32 131072 2>&1 | tail -1 | tee -a gpurun_out/compact_ab.log
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
