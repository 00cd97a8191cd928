timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-global --steps 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['kernel_ms_per_step'], 'fp64', d['fp64']['frac'])
print({k: (v['device_ms'], v['e2e_ms'], v['kernel_ms']) for k, v in d['single_window'].items()})
"
