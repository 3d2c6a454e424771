python -m pytest tests/test_env_gpu.py tests/test_edges_gpu.py tests/test_rollout_gpu.py -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-rollout --no-e2e --steps 100 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(b['ms_per_step'], b['roofline']['frac'], b['step_without_shaping'], b['expand4']['ms_per_launch'], b['expand4']['ms_per_launch_on_post_step_boards'])"
