# policy epilogue pass 2: 8-column chunks with the per-column vectors loaded one chunk ahead (variant) against 16-column chunks (main)
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/policy_pipe8_ab.log
for lib in "" $V/libhlynr_b200_pipe8.so; do
  HLYNR_B200_LIB=$lib timeout 600 python -m pytest tests/test_policy.py -m gpu -q -x 2>&1 | tail -1 | tee -a gpurun_out/policy_pipe8_ab.log
  HLYNR_B200_LIB=$lib timeout 300 python tools/policy_time.py 2>&1 | grep "n=131072\|n=1048576" | cut -c1-110 | tee -a gpurun_out/policy_pipe8_ab.log
  HLYNR_B200_LIB=$lib timeout 300 python tools/policy_phases.py 2>&1 | grep "cluster 1" | tee -a gpurun_out/policy_pipe8_ab.log
done
