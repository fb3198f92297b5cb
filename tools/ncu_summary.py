"""Summarise an .ncu-rep (read here, no GPU needed) and a launch-list CSV into profiles/.
usage: python tools/ncu_summary.py <prof.ncu-rep> <out.json> [launches.csv]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_tma_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    v = float(v)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for k, m in mult.items():
        if unit.startswith(k):
            return v * m
    return v


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:160]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": r[i], "unit": units[i]}
        st = sorted(((hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i] or 0))
                     for i in range(len(hdr)) if hdr[i].startswith("smsp__average_warps_issue_stalled_") and hdr[i].endswith("_per_issue_active.ratio")),
                    key=lambda x: -x[1])[:6]
        d["top_stalls_per_issue"] = st
        rd = to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"])
        wr = to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"])
        d["dram_bytes_per_launch"] = rd + wr
        res.append(d)
    summary = {"report": rep, "launches": res}
    if len(sys.argv) > 3:
        rows = [r for r in csv.reader(open(sys.argv[3])) if len(r) > 10]
        hi = rows[0].index("Metric Value")
        ki = rows[0].index("Kernel Name")
        tot = {}
        for r in rows[1:]:
            name = r[ki].split("(")[0][:80]
            t = tot.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += float(r[hi].replace(",", ""))
        unit = rows[1][rows[0].index("Metric Unit")]
        allt = sum(v[1] for v in tot.values())
        summary["launch_list"] = {"unit": unit, "kernels": {k: {"launches": v[0], "time": v[1], "share": v[1] / allt} for k, v in tot.items()}}
    json.dump(summary, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:90], d["gpu__time_duration.sum"], "dram B/launch", d["dram_bytes_per_launch"], d["top_stalls_per_issue"][:3])
    if "launch_list" in summary:
        for k, v in summary["launch_list"]["kernels"].items():
            print(f"{v['share']*100:6.2f}%  {v['launches']:4d}  {k}")


if __name__ == "__main__":
    main()
