#!/bin/bash
# hash-build path selection experiment: per-configuration stage times with the dense cell table forced off (0), at the default
# (1M cells) and at 4M cells
for m in 0 1048576 4194304; do
  echo "CIPC_DENSE_MIN_CELLS=$m"
  for c in cfg1 cfg2 cfg3 cfg4_50k cfg4_500k; do
    CIPC_DENSE_MIN_CELLS=$m python tools/stage_times.py $c 2>&1 | grep "^$c\|^{'ccs_hash_build\|hash_cells" | cut -c1-340
  done
done
