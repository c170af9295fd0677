#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_fold_round -c 1 -s 3 -f -o gpurun_out/prof_bls python scripts/kbench.py --vars 24 --iters 1 --modulus 52435875175126190479447740508185965837690552500527637822603658699938581184513 > gpurun_out/prof_bls.log 2>&1
tail -2 gpurun_out/prof_bls.log
