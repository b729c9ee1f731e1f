# One bench.py run summarised on one line: frame-iters/s, backward / raster ms, end to end.  Usage (under gpurun):
#   [ENV=...] bash tools/bench_line.sh LABEL [bench.py arguments]
label="$1"; shift
python bench.py --steps ${SWEEP_STEPS:-20} --warmup ${SWEEP_WARMUP:-5} --no-cpu-baseline --no-secondary "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); k = d['roofline']['kernel_ms_all']
print('$label', round(d['value']), 'ms', round(d['ms_per_step'], 4), ' '.join(f'{n} {v:.4f}' for n, v in k.items()), 'e2e', round(d['e2e']['value']))"
