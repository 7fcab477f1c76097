#!/bin/bash
mkdir -p gpurun_out
python tools/h5_device_decode_rate.py --nt 20 --out gpurun_out/c19_decode.json > gpurun_out/c19_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 2 -c 1 -f -o gpurun_out/c19_inflate python tools/h5_device_decode_rate.py --nt 20 --out gpurun_out/c19_decode_ncu.json > gpurun_out/c19_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/c19_plain.log
