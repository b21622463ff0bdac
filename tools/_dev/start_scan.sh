cd $GRAFT_REPO_ROOT
for v in "5e-5,1,1e-3,0.1,1e-9,4e-2" "5e-5,1,1e-3,0.1,1e-9,5e-2" "5e-5,1,1e-3,0.1,1e-9,7e-2"; do
  echo "== $v"
  GUSTO_IPM_MU0=$v python tools/_dev/hard_maxes.py 2>&1 | tail -1 | cut -c1-200
  GUSTO_IPM_MU0=$v python bench.py --no-cpu-baseline --no-extras --no-e2e --steps 90 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('easy bench', round(d['value']), d['ms_per_step'])"
done
