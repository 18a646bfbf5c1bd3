for E in 7104 14208 16384 21312 28416 65536 262144; do
python bench.py --no-cpu --steps 1000 --warmup 20 --e2e-steps 2 --envs $E 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print($E, 'fused us', round(d['breakdown_ms']['fused_step_obs']*1000,2), 'obs us', round(d['breakdown_ms']['observe_only']*1000,2), 'frac', round(d['roofline']['frac'],4), 'obs GB/s', round(d['breakdown_ms']['observe_only_GBps']))
"
done
