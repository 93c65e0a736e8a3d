#!/bin/bash
# occupancy sweep of the slice kernel (tuning builds): warps/SM = 4 * blocks_per_sm for 128-thread CTAs
for lib in t128_r124 t128_r80 t128_r64; do
  for bps in 1 2 3 4 5 6 8; do
    echo -n "bps=$bps "; B200SV_DEBUG_BLOCKS_PER_SM=$bps B200SV_LIB=$PWD/stochvolmodels_b200/lib/variants/libb200sv_$lib.so python tools/time_slice.py 2e7 136 0 2>&1 | grep libb200sv
  done
done
