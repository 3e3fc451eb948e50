#!/bin/bash
# Times bench.py stage breakdowns for every build_variants/*.so (SGR_LIB_PATH) — kernel experiment helper.
mkdir -p gpurun_out/var
for so in build_variants/*.so; do
  n=$(basename $so .so)
  SGR_LIB_PATH=$PWD/$so timeout 200 python bench.py --no-cpu-baseline --steps 60 --warmup 5 > gpurun_out/var/$n.json 2> gpurun_out/var/$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/var/*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], round(d["ms_per_step"], 4), {k: round(v["ms_per_step"] * 1e3) for k, v in d["roofline"]["stages"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
