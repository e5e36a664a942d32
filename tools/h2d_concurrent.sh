#!/bin/bash
# aggregate pinned host->device rate with 1, 2, 4, 8 concurrent processes (one GPU each):  tools/h2d_concurrent.sh > profiles/microbench/h2d_multi_b200.txt
BIN=$(dirname "$0")/../profiles/microbench/h2d_multi
NGPU=$(nvidia-smi -L | wc -l)
for P in 1 2 4 8; do
  [ "$P" -gt "$NGPU" ] && break
  START=$(python3 -c "import time; print(time.time() + 6)")
  for i in $(seq 0 $((P-1))); do "$BIN" "$i" "$START" 3 > /tmp/h2d_$i.txt & done
  wait
  echo "processes $P: per GPU $(for i in $(seq 0 $((P-1))); do awk '{printf "%s ", $2}' /tmp/h2d_$i.txt; done) GB/s, aggregate $(cat $(for i in $(seq 0 $((P-1))); do echo /tmp/h2d_$i.txt; done) | awk '{s+=$2} END {printf "%.1f", s}') GB/s"
done
