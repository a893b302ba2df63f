mkdir -p gpurun_out
python tools/aged_time.py cfg4,cfg2,cfg3 1 2>&1 | tee -a gpurun_out/cold_split.log
python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "4096 or sharding or statistics" 2>&1 | tail -3
