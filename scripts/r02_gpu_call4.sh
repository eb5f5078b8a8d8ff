#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
( time python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | tail -3
tail -3 gpurun_out/r02_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] if d["e2e"] else None)
print("roofline", {k: d["roofline"][k] for k in ("frac","traffic","frac_inside_timed_region","frac_whole_step","chain_ms_per_batch")})
print("per_frame", d.get("per_frame_api"))
print("dense", {k: d["dense_regime"].get(k) for k in ("value","speedup_vs_cpu")} if "dense_regime" in d else None)
for k,v in d.get("configs",{}).items(): print(k, v.get("frames_per_s"), v.get("chain_ms_per_call"), v.get("roofline_frac"), v.get("error"))
for k,v in d.get("next_rows",{}).items(): print(k, {kk: v.get(kk) for kk in ("frames_per_s","error","h2d_gbs","ms_per_clip")}, v.get("roofline",{}).get("frac"), v.get("cpu_baseline",{}).get("value"))
print("cpu", d.get("cpu_baseline",{}).get("value"))
PY
