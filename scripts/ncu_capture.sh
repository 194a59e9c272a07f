#!/bin/bash
# One full ncu capture of a kernel on the GPU box, summarised there (the .ncu-rep stays on the box:
# gpurun only brings 64 MiB back).   usage: ncu_capture.sh <name> <kernel regex> <skip> <which> <title> -- <command...>
name=$1; regex=$2; skip=$3; which=$4; title=$5; shift 6
timeout 400 ncu --set full --clock-control none -k regex:"$regex" -s "$skip" -c $((which + 1)) -o /tmp/$name "$@" > /tmp/$name.log 2>&1
python scripts/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/${name}_ncu_full.txt "$title" "$which" > /dev/null 2>> /tmp/$name.log || tail -5 /tmp/$name.log
