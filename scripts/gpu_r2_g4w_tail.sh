BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
for m in 22 20 18 16; do
SCB_PERSIST_MAX_GENERIC=$m timeout 600 python bench.py --modulus $BLS --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2w_bls_pmg$m.json 2> gpurun_out/r2w_bls_pmg$m.err
python - $m <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2w_bls_pmg%s.json"%sys.argv[1]))
print("persist_max_generic",sys.argv[1],"ms/proof",round(d["ms_per_step"],3),"verified",d.get("verified"),"launches",d.get("gpu_launches"))
PY
done
