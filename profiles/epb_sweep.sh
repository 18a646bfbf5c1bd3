for E in 2048 4096 8192; do for epb in 1 2 3 4; do echo -n "E=$E epb=$epb: "; MAPF_B200_EPB=$epb python profiles/rollout_probe.py c3 --envs $E 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('graph %.2f us (%.3f) rollout %.2f us (%.3f)'%(d['graph_us_per_step'],d['graph_frac_hbm'],d['rollout_us_per_step'],d['rollout_frac_hbm']))"; done; done
