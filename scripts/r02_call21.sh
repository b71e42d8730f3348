#!/bin/bash
# round 2, GPU call 21 (last ~2 minutes of box time): launch list of a full-size step of the final build (kernel shares for profiles/).
O=gpurun_out/call21; mkdir -p $O
timeout -k 5 95 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --min-seconds 0 > $O/launches_c3.log 2>&1
wc -l $O/launches_c3.csv
