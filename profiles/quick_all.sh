for W in c2 c3 c4; do
python bench.py --no-cpu --steps 1000 --warmup 20 --e2e-steps 2 --workload $W 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$W', 'fused us', round(d['breakdown_ms']['fused_step_obs']*1000,2), 'obs us', round(d['breakdown_ms']['observe_only']*1000,2), 'step us', round(d['breakdown_ms']['step_only']*1000,2), 'frac', round(d['roofline']['frac'],4), 'value', '%.3e' % d['value'], 'bfs ms', round(d['breakdown_ms']['goal_bfs_all_maps'],3))
"
done
python bench.py --no-cpu --steps 500 --warmup 20 --e2e-steps 2 --f32 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('c3 f32', 'fused us', round(d['breakdown_ms']['fused_step_obs']*1000,2), 'frac', round(d['roofline']['frac'],4))
"
