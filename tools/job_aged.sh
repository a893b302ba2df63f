mkdir -p gpurun_out
python tools/tick_series.py cfg4 > gpurun_out/series_cfg4.log 2>&1; tail -40 gpurun_out/series_cfg4.log
python tools/aged_step.py cfg4 > gpurun_out/aged.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/prof_aged python tools/aged_step.py cfg4 > gpurun_out/ncu_aged.log 2>&1
tail -3 gpurun_out/aged.log gpurun_out/ncu_aged.log
