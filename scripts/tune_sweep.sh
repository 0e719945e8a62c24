#!/bin/bash
# sweep the plane-sweep kernel's tile parameters on one operator (run under gpurun)
kind=${1:-heat}; nx=${2:-512}; ny=${3:-512}; nz=${4:-512}; bc=${5:-all}
for ys in 2 4; do for nt in 192 256 384; do for zc in 16 32 64 128; do for tx in 128 192 254; do
  PDE_B200_SW_YS=$ys PDE_B200_SW_NT=$nt PDE_B200_SW_ZC=$zc PDE_B200_SW_TXMAX=$tx timeout 120 python scripts/op_bench.py $kind $nx $ny $nz --bc $bc || echo "{\"fail\": \"$ys $nt $zc $tx\"}"
done; done; done; done
