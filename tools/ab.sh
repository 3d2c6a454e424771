python -m pytest tests/test_env_gpu.py tests/test_edges_gpu.py tests/test_rollout_gpu.py -x -q 2>&1 | tail -3
python tools/time_env.py 2>&1 | grep "step+shaping"
