"""Summarise an `ncu --page source --print-source cuda,sass --csv` export: per CUDA source line, the executed
warp-instructions and stall samples (lines with an Address of '-' are the per-line aggregates)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, agg = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) >= 8 and r[2] == "-":
            agg.append((cur_file, int(r[0]), r[1], float(r[hdr.index("# Samples")] or 0),
                        float(r[hdr.index("Instructions Executed")] or 0)))
    ti, ts = sum(a[4] for a in agg), sum(a[3] for a in agg)
    print(f"total warp-instructions {ti:.0f}, samples {ts:.0f}")
    for f, ln, src, smp, ins in sorted(agg, key=lambda a: -a[3])[:top]:
        print(f"{f}:{ln:<5d} inst {100 * ins / ti:5.1f}%  samples {100 * smp / ts:5.1f}%  {src.strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
