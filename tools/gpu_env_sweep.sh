# usage: bash tools/gpu_env_sweep.sh VAR v1 v2 ...   -- short bench per value of an environment switch, dw/tfar stage times
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 > gpurun_out/sweep_$v.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_$v.json"))
st=d["stages"]
print("$var=$v", "value", round(d["value"],1), " ".join(f"{k}={st[k]['ms_per_launch']:.4f}" for k in ("dw_s1","dw_s2_pool","tfar_global","tfar_cat_global","tfar_cat_local","dprnn_fused","resid_out","gate_proj")))
PY
done
