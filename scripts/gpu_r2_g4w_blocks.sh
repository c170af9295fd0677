BLS=52435875175126190479447740508185965837690552500527637822603658699938581184513
for b in 1 2; do
SCB_G4_BLOCKS=$b timeout 600 python bench.py --modulus $BLS --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2w_bls_blocks$b.json 2> gpurun_out/r2w_bls_blocks$b.err
python - $b <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2w_bls_blocks%s.json"%sys.argv[1]))
print("g4_blocks",sys.argv[1],"ms/proof",round(d["ms_per_step"],2),"verified",d.get("verified"),"kernel_ms",round(d["roofline"]["kernel_ms"],2))
PY
done
