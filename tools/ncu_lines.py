"""Per-CUDA-line instruction attribution from an ncu report captured with --import-source on.

  python tools/ncu_lines.py gpurun_out/<rep>.ncu-rep [units] [top]
units = number of work units (e.g. 8x8 tile-candidates) to normalise thread-instructions by.
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, agg = None, []

    def I(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in csv.reader(raw.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 10 or r[0] == "Line No" or r[0] == "":
            continue
        agg.append((cur, I(r[0]), r[1].strip(), I(r[7]), I(r[8]), I(r[6])))
    tot = sum(a[3] for a in agg)
    tth = sum(a[4] for a in agg)
    print("warp instructions %d, thread instructions %d%s" % (tot, tth, ", per unit %.1f" % (tth / units) if units else ""))
    agg.sort(key=lambda a: -a[3])
    for a in agg[:top]:
        print("%-16s %4d %6.2f%% %s samp %6d  %s" % (a[0], a[1], 100.0 * a[3] / tot,
                                                    "(%6.1f/unit)" % (a[4] / units) if units else "", a[5], a[2][:100]))


if __name__ == "__main__":
    main()
