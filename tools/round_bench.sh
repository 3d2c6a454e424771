# full GPU test suite, the bench line, the reference arm, and the ncu evidence (launch list + full capture)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --rollout-envs 18944 --rollout-steps 32 --urm-envs 4736 --urm-steps 2 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel_dense -s 2 -c 1 -f -o gpurun_out/r01_step_dense python tools/run_step.py 4 > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
