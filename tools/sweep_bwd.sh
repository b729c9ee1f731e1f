# build variants of the backward kernel on the GPU box and bench each (run under gpurun from the repo root)
# each argument: "<nvcc -D flags>|<env assignments>"
run() {
  flags="${1%%|*}"; envs="${1#*|}"; [ "$envs" = "$1" ] && envs=""
  DH_EXTRA_NVCC_FLAGS="$flags" python -m dynhor_b200.build --force > /dev/null 2>&1
  env $envs python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline']['kernel_ms_all']['backward'], d['roofline']['kernel_ms_all']['raster'])"
}
for v in "$@"; do run "$v"; done
python -m dynhor_b200.build --force > /dev/null 2>&1
