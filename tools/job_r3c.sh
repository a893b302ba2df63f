set -x
mkdir -p gpurun_out
HLYNR_PRECISION=fp64 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_aged_f64 python tools/aged_step.py cfg4 > gpurun_out/ncu_aged_f64.log 2>&1
tail -2 gpurun_out/ncu_aged_f64.log
