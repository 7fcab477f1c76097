#!/bin/bash
# round-2 GPU call 3: ncu --set full of the fused pass on the shapes that matter (per-rank C4 / C5 shares, f64 and f32)
mkdir -p gpurun_out
run() {  # name workload nt dtype
  python tools/run_once.py --workload $2 --nt $3 --dtype $4 --passes 3 --out gpurun_out/c3_$1.json > gpurun_out/c3_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k23_fused -s 1 -c 1 -f -o gpurun_out/c3_$1 \
      python tools/run_once.py --workload $2 --nt $3 --dtype $4 --passes 1 > gpurun_out/c3_$1.ncu.log 2>&1
  echo "$1 rc=$?"
}
run C3_f32 C3 64 f32
run C3_f64 C3 64 f64
run C4_f64 C4 8 f64
run C4_f32 C4 8 f32
run C5_f64 C5 2 f64
run C5_f32 C5 2 f32
