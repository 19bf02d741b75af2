#!/bin/bash
# launch list of one detection (cold-cache, serialised) -> gpurun_out/$1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$1 python tools/one_detection.py > gpurun_out/ncu_one.log 2>&1
