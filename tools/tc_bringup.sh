#!/bin/bash
# bring-up runs of the tensor-core Activation1d harness (each bounded by timeout; failures do not stop the list)
H=tools/_build/tc_harness
mkdir -p gpurun_out
L=gpurun_out/tc_bringup.log
: > $L
run() { echo "=== $*" >> $L; timeout 90 $H "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run 1 128 256 --ny 4 --debug 1
run 1 128 256 --ny 4 --debug 2
run 2 24 2048 --ny 4 --debug 1
run 1 128 4096 --ny 64 --debug 1
run 2 24 8192 --ny 32 --debug 2
run 3 40 1000 --ny 64
run 3 40 1000 --ny 8
run 5 7 4104 --ny 20
run 2 24 220416 --iters 20
run 16 24 220416 --iters 20 --check-rows 8
run 16 48 110208 --iters 20 --check-rows 8
run 16 384 13776 --iters 20 --check-rows 8
run 16 768 3440 --iters 20 --check-rows 8
run 32 96 2048 --iters 20 --check-rows 16
run 2 512 8192 --iters 20 --check-rows 16
tail -c 6000 $L
