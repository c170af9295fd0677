#!/bin/bash
# round 2, third session: from which table size should the first pair pass be a launch of its own when the 21-bit triples apply?
set -u
mkdir -p gpurun_out
P=gpurun_out/r2i
rm -f ${P}_first_alone.jsonl
for v in 25 24 23; do for fa in 26 $v; do SCB_PAIR_FIRST_ALONE=$fa timeout 120 python scripts/kbench_w21.py $v 0,3,0,3 >> ${P}_first_alone.jsonl 2>> ${P}_first_alone.err; done; done
cat ${P}_first_alone.jsonl | cut -c1-260
