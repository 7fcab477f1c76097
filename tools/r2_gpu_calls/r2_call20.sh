#!/bin/bash
# round-2 validation on one B200: tests (plain and with guard bands), smoke, both bench arms as the driver runs them, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c20_pytest.log
NFX_DEBUG_GUARDS=1 python -m pytest tests -m gpu -x -q > gpurun_out/c20_pytest_guards.log 2>&1; echo "pytest with guard bands rc=$?"; tail -2 gpurun_out/c20_pytest_guards.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c20_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c20_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 --impl reference > gpurun_out/c20_ref_n1.json 2> gpurun_out/c20_ref_n1.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c20_bench_n1.json 2> gpurun_out/c20_bench_n1.err; echo "bench rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 --dtype f32 > gpurun_out/c20_bench_n1_f32.json 2> gpurun_out/c20_bench_n1_f32.err; echo "bench f32 rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/c20_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c20_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/c20_ncu.log 2>&1
echo "launch list rc=$?"
