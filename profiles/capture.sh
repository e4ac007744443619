#!/bin/bash
# ncu evidence of one bench workload (run under gpurun): launch list of the whole bench command, then one --set full
# capture of the dominant kernel.  usage: bash profiles/capture.sh <config> <kernel regex> <tag>
set -e
CFG=$1; PAT=$2; TAG=$3
CMD="python bench.py --config $CFG --steps 12 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --roofline-ms 20 --prewarm-ms 50"
$CMD > gpurun_out/${TAG}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$PAT -s 24 -c 2 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "$TAG done"
