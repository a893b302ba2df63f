set -x
mkdir -p gpurun_out
HLYNR_OPTS=split=1 ncu --set full --clock-control none --import-source on -k regex:step_kernel_ws -s 5 -c 1 -f -o gpurun_out/prof_ws python tools/aged_step.py cfg4 > gpurun_out/ncu_ws.log 2>&1
tail -3 gpurun_out/ncu_ws.log
