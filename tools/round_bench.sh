# Round evidence in one gpurun call: full GPU test suite, the bench line, the reference arm, the launch list and the full ncu
# captures of the headline kernels (each only after its own command has exited 0 without ncu).
set -x
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --rollout-envs 18944 --rollout-steps 32 --c4-envs 0 --urm-envs 4736 --urm-steps 2 --urm-train-envs 1024"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python tools/run_step4.py 4 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step4_kernel_dense -s 2 -c 1 -f -o gpurun_out/step4 python tools/run_step4.py 4 > gpurun_out/ncu_step4.log 2>&1; tail -2 gpurun_out/ncu_step4.log
python tools/run_step.py 4 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel_dense -s 2 -c 1 -f -o gpurun_out/step_dense python tools/run_step.py 4 > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
