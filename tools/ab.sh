# A/B of kernel variants on the GPU box (run under gpurun from the repo root): for every argument
# "<nvcc -D flags>|<env assignments>" rebuild dh_jointopt.cu with the flags (the default build is restored at the end)
# and print tools/bench_line.sh's line.      SWEEP_STEPS=30 bash tools/ab.sh "" "-DDH_FLAT_ENUM=0" "|DH_BWD_CHUNKS=12"
for v in "$@"; do
  flags="${v%%|*}"; envs="${v#*|}"; [ "$envs" = "$v" ] && envs=""
  DH_EXTRA_NVCC_FLAGS="$flags" python -m dynhor_b200.build dh_jointopt.cu > /dev/null 2>&1 || { echo "$v BUILD FAILED"; continue; }
  env $envs bash tools/bench_line.sh "[$v]"
done
python -m dynhor_b200.build dh_jointopt.cu > /dev/null 2>&1
