#!/bin/bash
# A/B of the bottleneck-tail kernel's shared-memory split (ring slots vs output staging buffers): whole plan + layer1/2
cd "$(dirname "$0")/.."
for cfg in "1 0" "1 2" "1 3" "0 3" "0 4" "0 2"; do
  set -- $cfg
  echo "== UP_TAIL_TALL=$1 UP_TAIL_OBUFS=$2 (0 = default)"
  UP_TAIL_TALL=$1 UP_TAIL_OBUFS=$2 timeout 60 python tools/segment_bench.py 2>&1 | grep -E "whole plan|layer1|layer2"
done
