"""Copies the evidence of tools/capture_profiles.sh (gpurun_out/prof/) into profiles/ under round-tagged names and
derives profiles/<round>_traffic.json (DRAM bytes per launch of the blend kernels from the `ncu --set full` capture),
which bench.py reports as roofline.traffic."""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out", "prof")
dst = os.path.join(ROOT, "profiles")
for a, b in (("bench.json", "bench.json"), ("bench_reference.json", "bench_reference.json"), ("launches.csv", "launches.csv"),
             ("blend_details.txt", "blend_details.txt"), ("small_details.txt", "small_details.txt"), ("prep_details.txt", "prep_details.txt"),
             ("blend_raw.csv", "blend_raw.csv")):
    if os.path.exists(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, f"{tag}_{b}"))
rows = list(csv.reader(open(os.path.join(src, "blend_raw.csv"))))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"source": f"profiles/{tag}_blend_raw.csv (ncu --set full --clock-control none, one launch each, cold caches)"}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    key = "blend_forward" if "blend_forward" in name else "blend_backward" if "blend_backward" in name else None
    if key is None:
        continue
    rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
    out[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")]),
                "warp_instructions": float(r[hdr.index("smsp__inst_executed.sum")])}
json.dump(out, open(os.path.join(dst, f"{tag}_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
