"""Per-source-line summary of an ncu capture: `python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]`.
Aggregates the source page (needs -lineinfo + --import-source on) to instructions executed and stall samples per
CUDA source line, so that the cost of each phase of a kernel can be read off."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                          "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or r[0] == "":
            continue
        lines.append(r)
    i_inst = hdr.index("Instructions Executed")
    i_samp = hdr.index("# Samples")
    tot_i = sum(int(r[i_inst]) for r in lines)
    tot_s = sum(int(r[i_samp]) for r in lines)
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    lines.sort(key=lambda r: -int(r[i_inst]))
    for r in lines[:top]:
        print(f"{r[0]:>5} inst {100 * int(r[i_inst]) / tot_i:5.1f}%  samp {100 * int(r[i_samp]) / max(tot_s, 1):5.1f}%  {r[1][:110]}")


if __name__ == "__main__":
    main()
