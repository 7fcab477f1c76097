#!/bin/bash
mkdir -p gpurun_out
NFX_DEBUG_GUARDS=1 python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest_guards.log 2>&1; echo "pytest with guard bands rc=$?"; tail -3 gpurun_out/c14_pytest_guards.log
python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c14_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/c14_bench_n1_f64.json 2> gpurun_out/c14_bench_n1_f64.err; echo "bench f64 rc=$?"
python bench.py --steps 20 --warmup 5 --dtype f32 > gpurun_out/c14_bench_n1_f32.json 2> gpurun_out/c14_bench_n1_f32.err; echo "bench f32 rc=$?"
python bench.py --steps 5 --warmup 3 --impl reference > gpurun_out/c14_ref_n1.json 2> gpurun_out/c14_ref_n1.err; echo "ref rc=$?"
run() {  # name workload nt dtype
  python tools/run_once.py --workload $2 --nt $3 --dtype $4 --passes 3 --out gpurun_out/c14_$1.json > gpurun_out/c14_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k23_fused -s 1 -c 1 -f -o gpurun_out/c14_$1 \
      python tools/run_once.py --workload $2 --nt $3 --dtype $4 --passes 1 > gpurun_out/c14_$1.ncu.log 2>&1
  echo "$1 rc=$?"
}
run C3_f64 C3 64 f64
run C3_f32 C3 64 f32
run C4_f64 C4 8 f64
run C4_f32 C4 8 f32
run C5_f64 C5 2 f64
run C5_f32 C5 2 f32
