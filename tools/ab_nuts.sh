#!/usr/bin/env bash
# A/B of the sampler variants on one box.  Writes gpurun_out/$AB_OUT
out=gpurun_out/${AB_OUT:-r02_ab_nuts.log}
: > $out
export QB_NUTS_ONLY=1 QB_NPROB=${QB_NPROB:-1776} QB_ITER=${QB_ITER:-200}
run() { echo "== $*" >> $out; env "$@" python tools/quick_bench.py 2>&1 | grep -E "NUTS|Error|error" >> $out; }
for v in "$@"; do run $v; done
cat $out
