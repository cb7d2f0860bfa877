"""Rank the CUDA source lines of a kernel in an .ncu-rep by executed warp instructions and stall samples.
python tools/ncu_lines.py report.ncu-rep [top=40]      (runs in the build container: `ncu -i` needs no GPU)"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("", "Function Name") or not r[0].isdigit():
            continue
        ix = hdr.index("Instructions Executed")
        sx = hdr.index("# Samples")
        try:
            lines.append((int(r[ix]), int(r[sx]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print(f"total warp instructions {tot_i:.3e}, stall samples {tot_s}")
    print("  inst%  samp%  file:line  source")
    for n, s, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"  {100.0 * n / tot_i:5.1f}  {100.0 * s / tot_s:5.1f}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main()
