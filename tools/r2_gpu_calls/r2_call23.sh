#!/bin/bash
# final 1-GPU validation of the round: whole GPU suite (plain + guard bands), smoke, bench as the driver runs it, launch list of the timed steps
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c23_pytest.log
NFX_DEBUG_GUARDS=1 python -m pytest tests -m gpu -x -q > gpurun_out/c23_pytest_guards.log 2>&1; echo "pytest with guard bands rc=$?"; tail -2 gpurun_out/c23_pytest_guards.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c23_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c23_bench_n1.json 2> gpurun_out/c23_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/c23_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k23_fused|k_reduce_subrows|k_or_flag|k_probe_read' -c 60 --csv --log-file gpurun_out/c23_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/c23_ncu.log 2>&1
echo "launch list rc=$?"
