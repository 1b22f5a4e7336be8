"""Summarise ncu reports brought back from the GPU box into small tracked files under profiles/.

    python bench_tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches
    python bench_tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  profiles/r1_group_fwd  [json-out]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].split("(")[0][:80]
        tot[name][0] += 1
        tot[name][1] += float(r["Metric Value"].replace(",", ""))
    total = sum(v[1] for v in tot.values())
    with open(dst + "_summary.md", "w") as o:
        o.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
            o.write("| `%s` | %d | %.1f | %.2f | %.3f |\n" % (k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, v[1] / total))
    with open(dst + ".csv", "w") as o:
        o.writelines(lines)
    print(open(dst + "_summary.md").read())


def kernel(src, dst, json_out=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        out.append({"kernel": d.get("Kernel Name", "")[:100], **{k: d[k] for k in KEYS if k in d}})
    with open(dst + ".json", "w") as o:
        json.dump(out, o, indent=1)
    print(json.dumps(out, indent=1)[:3000])
    if json_out and out:
        k = out[0]
        rd = float(k.get("dram__bytes_read.sum", "0").replace(",", ""))
        wr = float(k.get("dram__bytes_write.sum", "0").replace(",", ""))
        # ncu prints dram bytes in the unit of the column header row (rows[1]); normalise to bytes
        unit = dict(zip(hdr, rows[1])).get("dram__bytes_read.sum", "byte")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        with open(json_out, "w") as o:
            json.dump({"kernel": k["kernel"], "dram_bytes_per_launch": (rd + wr) * mult,
                       "dram_read_bytes": rd * mult, "dram_write_bytes": wr * mult,
                       "source": src.split("/")[-1], "how": "ncu --set full --clock-control none"}, o, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
