set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
HLYNR_B200_LIB=$V/libhlynr_b200_epi16.so ncu --set full --clock-control none --import-source on -k regex:policy_forward -s 3 -c 1 -f -o gpurun_out/prof_policy16 python tools/policy_step.py > gpurun_out/ncu_policy16.log 2>&1
tail -3 gpurun_out/ncu_policy16.log
