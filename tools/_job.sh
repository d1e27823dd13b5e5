#!/bin/bash
mkdir -p gpurun_out
CONTRAST_STEPS=20 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_contrast_s38.csv python tools/contrast_bench.py yelp2018 10 none > gpurun_out/ncu_contrast_s38.log 2>&1
python tools/launch_summary.py gpurun_out/launches_contrast_s38.csv 2>/dev/null | head -22
