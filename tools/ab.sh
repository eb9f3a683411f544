#!/bin/bash
# A/B timing on the GPU box: tools/ab.sh WORKLOAD name1 name2 ...  (names of recoup_b200/variants/lib_NAME.so,
# "base" = the in-tree library).  One short device-resident bench per library; prints the stage table.
wl=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset RCP_LIB_PATH; else export RCP_LIB_PATH=$PWD/recoup_b200/variants/lib_$v.so; fi
  python bench.py --workload $wl --configs none --strong off --no-cpu-baseline --steps 10 --warmup 3 \
      > gpurun_out/ab_${wl}_$v.json 2> gpurun_out/ab_${wl}_$v.err || echo "$v FAILED: $(tail -n 3 gpurun_out/ab_${wl}_$v.err)"
  python - "$v" gpurun_out/ab_${wl}_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    st = {k: round(v, 4) for k, v in d["stage_ms_per_step"].items()}
    fu = d.get("fused") or {}
    print("%-14s %.4f ms/step  cov frac %.3f  fused %.4f  %s" % (
        sys.argv[1], d["ms_per_step"], d["roofline"]["stages"]["coverage"]["frac"],
        fu.get("ms_per_step", 0.0), st))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
