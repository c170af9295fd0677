#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (subset with 4-limb fields)"
timeout 1500 python -m pytest tests -m gpu -x -q -k "p255 or golden or triangle or gkr or mle or large" 2>&1 | tail -6
echo "== kbench BLS12-381 Fr"
python scripts/kbench.py --vars 26 --iters 5 --modulus 52435875175126190479447740508185965837690552500527637822603658699938581184513
python scripts/kbench.py --vars 26 --iters 5 --tables 2 --modulus 52435875175126190479447740508185965837690552500527637822603658699938581184513
echo "== kbench Goldilocks"
python scripts/kbench.py --vars 28 --iters 5 --modulus 18446744069414584321
