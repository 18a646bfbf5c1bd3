#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "primal" 2>&1 | tail -4
python bench.py --no-cpu --steps 2000 --warmup 20 --e2e-steps 2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('fused us', d['breakdown_ms']['fused_step_obs']*1000, 'obs us', d['breakdown_ms']['observe_only']*1000, 'step us', d['breakdown_ms']['step_only']*1000, 'frac', d['roofline']['frac'], 'clk', d['clocks'])
"
