set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_aged_cfg3 python tools/aged_step.py cfg3 > gpurun_out/ncu_aged_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -f -o gpurun_out/prof_aged_cfg2 python tools/aged_step.py cfg2 > gpurun_out/ncu_aged_cfg2.log 2>&1
tail -1 gpurun_out/ncu_aged_cfg3.log gpurun_out/ncu_aged_cfg2.log
