"""Per-kernel time and DRAM traffic of the LAST step of an ncu launch list captured with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv ...
usage: ncu_traffic.py <csv> <launches per step> [out.json]   (writes profiles/r02_bench_step_traffic.json-style summary)"""
import csv
import json
import re
import sys


def _bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(path, per_step, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    by_id = {}
    order = []
    for x in csv.DictReader(lines):
        i = int(x["ID"])
        if i not in by_id:
            by_id[i] = {"name": re.sub(r"\(.*", "", x["Kernel Name"]).split("<")[0].split("::")[-1].replace("void ", "")}
            order.append(i)
        m = x["Metric Name"]
        if m == "gpu__time_duration.sum":
            u = x["Metric Unit"]
            by_id[i]["ns"] = float(x["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)
        elif m == "dram__bytes_read.sum":
            by_id[i]["rd"] = _bytes(x["Metric Value"], x["Metric Unit"])
        elif m == "dram__bytes_write.sum":
            by_id[i]["wr"] = _bytes(x["Metric Value"], x["Metric Unit"])
    last = [by_id[i] for i in order[-per_step:]]
    agg = {}
    for r in last:
        a = agg.setdefault(r["name"], {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        a["launches"] += 1
        a["ms"] += r.get("ns", 0) / 1e6
        a["dram_read_bytes"] += r.get("rd", 0)
        a["dram_write_bytes"] += r.get("wr", 0)
    tot = sum(a["ms"] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        a["share"] = a["ms"] / tot
        print("%-34s n=%4d %8.2f ms %5.1f%%  rd %8.1f MB  wr %8.1f MB  %6.2f TB/s" % (
            k[:34], a["launches"], a["ms"], 100 * a["share"], a["dram_read_bytes"] / 1e6, a["dram_write_bytes"] / 1e6,
            (a["dram_read_bytes"] + a["dram_write_bytes"]) / 1e12 / max(a["ms"] / 1e3, 1e-12)))
    print("total %.2f ms over %d launches" % (tot, len(last)))
    if out:
        g = agg.get("gemm_tc_kernel", {})
        json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control "
                             "none over bench.py (last step of %s, %d launches)" % (path, len(last)),
                   "gemm_tc_kernel_dram_bytes_per_step": g.get("dram_read_bytes", 0) + g.get("dram_write_bytes", 0),
                   "step_dram_bytes": sum(a["dram_read_bytes"] + a["dram_write_bytes"] for a in agg.values()),
                   "by_kernel": agg}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else None)
