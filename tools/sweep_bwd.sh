for t in 256 288; do DH_EXTRA_NVCC_FLAGS="-DDH_BWD_THREADS=$t" python -m dynhor_b200.build --force > /dev/null 2>&1; python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bwd threads $t', d['value'], d['roofline']['kernel_ms_all']['backward'])"; done
python -m dynhor_b200.build --force > /dev/null 2>&1
python -m pytest tests -m gpu -q -k "not 20k" 2>&1 | tail -2
