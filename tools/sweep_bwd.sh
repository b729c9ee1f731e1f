# Build kernel variants on the GPU box and bench each (run under gpurun from the repo root); the default build is
# restored at the end.  Each argument: "<nvcc -D flags>|<env assignments>", e.g.
#   bash tools/sweep_bwd.sh "" "-DDH_PAIR_CAP=12" "-DDH_BWD_THREADS=256 -DDH_BWD_MIN_CTAS=3" "|DH_BWD_CHUNKS=12" \
#        "-DDH_DEFER_DEPTH=1" "-DDH_TILE_Z=0" "-DDH_STRIP_ROWS=8 -DDH_RASTER_THREADS=256 -DDH_RASTER_MIN_CTAS=3"
# Knobs (dh_jointopt.cu): DH_BWD_THREADS, DH_BWD_MIN_CTAS, DH_CHUNK_FACES, DH_PAIR_CAP, DH_LISTS_GLOBAL, DH_EVEN_LAST,
# DH_FAST_COEF, DH_FIDX_NOALLOC, DH_RASTER_THREADS, DH_RASTER_MIN_CTAS, DH_STRIP_ROWS, DH_RASTER_EVEN,
# DH_RASTER_SPLIT, DH_TILE_Z, DH_DEFER_DEPTH, DH_PASS_ORDER; env: DH_BWD_CHUNKS.
# Prints: flags, frame-iters/s, backward-segment ms, raster ms (bench.py --steps 100 --warmup 20).
run() {
  flags="${1%%|*}"; envs="${1#*|}"; [ "$envs" = "$1" ] && envs=""
  DH_EXTRA_NVCC_FLAGS="$flags" python -m dynhor_b200.build --force > /dev/null 2>&1
  env $envs python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline']['kernel_ms_all']['backward'], d['roofline']['kernel_ms_all']['raster'])"
}
for v in "$@"; do run "$v"; done
python -m dynhor_b200.build --force > /dev/null 2>&1
