#!/bin/bash
# k_corr tuning sweep (run on the GPU box): rebuild with different CTAs/SM x stages, kernel time from ncu
mkdir -p gpurun_out
for cfg in "3 3" "2 4" "2 3" "3 2" "4 2"; do
  set -- $cfg
  DH_EXTRA_NVCC_FLAGS="-DDH_CORR_CTAS=$1 -DDH_CORR_STAGES=$2" python -m dynhor_b200.build --force > /dev/null
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_corr --csv \
      --log-file gpurun_out/corr_sweep_$1_$2.csv python tools/bench_corr.py > gpurun_out/corr_sweep_$1_$2.log 2>&1
done
python -m dynhor_b200.build --force > /dev/null
