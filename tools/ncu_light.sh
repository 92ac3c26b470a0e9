#!/bin/bash
# usage: tools/ncu_light.sh out.csv kernel-regex command...   (a few counters of one launch: instruction caches, issue rate, stalls)
out=$1; k=$2; shift 2
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for s in no_instruction wait short_scoreboard long_scoreboard branch_resolving barrier membar math_pipe_throttle lg_throttle; do
  M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio
done
ncu --metrics $M --clock-control none -k regex:$k -s 1 -c 1 --csv --log-file $out "$@"
