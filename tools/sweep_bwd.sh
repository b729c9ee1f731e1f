# Build kernel variants on the GPU box and bench each (run under gpurun from the repo root); the default build is
# restored at the end.  Each argument: "<nvcc -D flags>|<env assignments>", e.g.
#   bash tools/sweep_bwd.sh "" "-DDH_PAIR_CAP=12" "-DDH_BWD_THREADS=256 -DDH_BWD_MIN_CTAS=3" "|DH_BWD_CHUNKS=12" \
#        "-DDH_DEFER_DEPTH=1" "-DDH_TILE_Z=0" "-DDH_STRIP_ROWS=8 -DDH_RASTER_THREADS=256 -DDH_RASTER_MIN_CTAS=3"
# Knobs (dh_jointopt.cu): DH_BWD_THREADS, DH_BWD_MIN_CTAS, DH_CHUNK_FACES, DH_PAIR_CAP, DH_LISTS_GLOBAL, DH_EVEN_LAST,
# DH_FAST_COEF, DH_FIDX_NOALLOC, DH_RASTER_THREADS, DH_RASTER_MIN_CTAS, DH_STRIP_ROWS, DH_RASTER_EVEN,
# DH_RASTER_SPLIT, DH_TILE_Z, DH_DEFER_DEPTH, DH_PASS_ORDER; env: DH_BWD_CHUNKS.
# Prints: flags, frame-iters/s, backward-segment ms, raster ms, end-to-end (bench.py --steps 30 --warmup 5; SWEEP_STEPS /
# SWEEP_WARMUP override).
run() {
  flags="${1%%|*}"; envs="${1#*|}"; [ "$envs" = "$1" ] && envs=""
  DH_EXTRA_NVCC_FLAGS="$flags" python -m dynhor_b200.build dh_jointopt.cu > /dev/null 2>&1 || { echo "$1 BUILD FAILED"; return; }
  env $envs python bench.py --steps ${SWEEP_STEPS:-30} --warmup ${SWEEP_WARMUP:-5} --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_all']; print('$1', round(d['value']), 'bwd', round(k['backward'],4), 'raster', round(k['raster'],4), 'e2e', round(d['e2e']['value']))"
}
for v in "$@"; do run "$v"; done
python -m dynhor_b200.build dh_jointopt.cu > /dev/null 2>&1
