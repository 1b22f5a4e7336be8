"""Per-source-line instruction and stall-sample shares from an ncu report (needs -lineinfo + --import-source on).

    python bench_tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import csv
import subprocess
import sys


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fname, agg, seen_fn = None, {}, 0
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            continue
        if len(r) < 10 or r[0] in ("Line No", ""):
            continue
        key = (fname, num(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += num(r[7])
        a[1] += num(r[6])
        a[2] += num(r[31])
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print("total warp instructions %d, stall samples %d (all profiled launches)" % (tot, ts))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5.1f%% inst %5.1f%% samples (barrier %5d) | %s:%d %s" % (100 * v[0] / tot, 100 * v[1] / ts, v[2], *k))


if __name__ == "__main__":
    main()
