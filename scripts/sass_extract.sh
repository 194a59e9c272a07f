#!/bin/bash
# SASS evidence for profiles/: which Blackwell-native instructions each hot kernel contains
# (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit).
so=safe-grid-agents_b200/gridfast/libsgk.so
out=${1:-profiles/r02_sass_summary.txt}
{
echo "cuobjdump -sass $so -- per-kernel counts of tensor-core / TMEM / TMA instructions ($(date -u +%F))"
echo "nvcc $(nvcc --version | grep release | sed 's/.*release //')"
for k in k_mlp_forward_ts k_mlp_backward_fused k_wgrad_mn k_mlp_backward_data_tc k_wgrad_tc "k_rollout_privateILi0E12PhiloxStreamLb0ELb0ELi1ELi0E" "k_rollout_privateILi1E12PhiloxStreamLb0ELb0ELi2E"; do
  for fn in $(cuobjdump -sass $so 2>/dev/null | grep "Function :" | grep "$k" | awk '{print $3}' | sort -u); do
    echo; echo "== $(echo $fn | c++filt)"
    cuobjdump -sass -fun "$fn" $so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | awk '{print $1}' | sed 's/;$//' > /tmp/sass_ops.txt
    echo "   instructions: $(wc -l < /tmp/sass_ops.txt)"
    grep -E "^(UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|HMMA|DADD|DMUL|DFMA|DSETP|LDS|STS|LDG|STG)" /tmp/sass_ops.txt | sort | uniq -c | sort -rn | awk '{printf "   %6d %s\n", $1, $2}'
  done
done
echo; echo "whole library: $(cuobjdump -sass $so 2>/dev/null | grep -c 'Function :') kernels; tcgen05.mma instructions: $(cuobjdump -sass $so 2>/dev/null | grep -c UTCHMMA)"
} > $out
