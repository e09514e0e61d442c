for d in 15 31 47 79 143 111 127 255; do SEPCORE_DEBUG_SKIP=$d python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --streams 1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip=$d', 'step_us %.1f' % (1e3*d['ms_per_step']), 'kernel_us %.1f' % (1e3*d['roofline']['kernel_ms']))"; done
