"""Summarise the CSV exports of an ncu capture (tools/ncu_job.sh: <name>.raw.csv + <name>.source.csv) into profiles/.
usage: python tools/ncu_csv_summary.py <prefix-without-.raw.csv> <out.json>"""
import collections
import csv
import json
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_summary import KEYS, to_bytes  # noqa: E402


def main():
    pre, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(pre + ".raw.csv")))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = {"kernel": r[hdr.index("Kernel Name")][:200]}
    for k in KEYS:
        if k in hdr:
            d[k] = {"value": r[hdr.index(k)], "unit": units[hdr.index(k)]}
    d["top_stalls_per_issue"] = sorted(((hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i] or 0))
                                        for i in range(len(hdr)) if hdr[i].startswith("smsp__average_warps_issue_stalled_") and hdr[i].endswith("_per_issue_active.ratio")),
                                       key=lambda x: -x[1])[:8]
    d["dram_bytes_per_launch"] = to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"]) + \
        to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"])
    try:
        src = [x for x in list(csv.reader(open(pre + ".source.csv")))[2:] if len(x) > 6]
        tot = sum(int(x[4]) for x in src)
        by = collections.Counter()
        for x in src:
            t = x[1].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            by[op] += int(x[4])
        d["sass"] = {"static_instructions": len(src), "warp_instructions_executed": sum(int(x[5]) for x in src), "stall_samples": tot,
                     "samples_by_opcode_pct": {k: round(100.0 * v / max(tot, 1), 1) for k, v in by.most_common(12)},
                     "hottest": [{"samples": int(x[4]), "executed": int(x[5]), "sass": x[1].strip()[:100]} for x in sorted(src, key=lambda x: -int(x[4]))[:12]]}
    except FileNotFoundError:
        pass
    json.dump({"source": pre, "launch": d}, open(out, "w"), indent=1)
    print(out, d["gpu__time_duration.sum"], d["dram_bytes_per_launch"])


if __name__ == "__main__":
    main()
