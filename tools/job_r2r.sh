# policy forward: 16 epilogue warps (4 per TMEM lane quarter) against 8
set -x
mkdir -p gpurun_out
V=$PWD/hlynr_intercept_b200/_variants
rm -f gpurun_out/policy_epi_ab.log
for lib in "" $V/libhlynr_b200_epi16.so; do
  HLYNR_B200_LIB=$lib timeout 600 python -m pytest tests/test_policy.py -m gpu -q -x 2>&1 | tail -2 | tee -a gpurun_out/policy_epi_ab.log
  HLYNR_B200_LIB=$lib timeout 300 python tools/policy_time.py 2>&1 | tail -8 | tee -a gpurun_out/policy_epi_ab.log
  HLYNR_B200_LIB=$lib timeout 300 python tools/policy_phases.py 2>&1 | tail -12 | tee -a gpurun_out/policy_epi_ab.log
done
