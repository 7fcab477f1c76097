#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_field_cli.py tests/test_gpu_inflate.py -x -q -m gpu > gpurun_out/c18_pytest.log 2>&1; tail -3 gpurun_out/c18_pytest.log
python tools/h5_device_decode_rate.py --nt 64 --out gpurun_out/c18_h5_device_decode.json > gpurun_out/c18_decode.log 2>&1; tail -8 gpurun_out/c18_decode.log
