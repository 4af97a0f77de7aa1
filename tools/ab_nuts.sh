#!/usr/bin/env bash
# A/B of the sampler variants on one box: runtime-flags kernel vs the compile-time specialised one (12 / 16 warps per
# SM, with / without the tick barrier).  Writes gpurun_out/r02_ab_nuts.log
out=gpurun_out/${AB_OUT:-r02_ab_nuts.log}
: > $out
export QB_NUTS_ONLY=1 QB_NPROB=${QB_NPROB:-1776} QB_ITER=${QB_ITER:-200}
run() { echo "== $*" >> $out; env "$@" python tools/quick_bench.py 2>&1 | grep -E "NUTS|Error|error" >> $out; }
run BSYN_NO_FAST=1
run BSYN_NUTS_NW=12 BSYN_NUTS_TICK=1
run BSYN_NUTS_NW=12 BSYN_NUTS_TICK=0
run BSYN_NUTS_NW=16 BSYN_NUTS_TICK=1
run BSYN_NUTS_NW=16 BSYN_NUTS_TICK=0
cat $out
