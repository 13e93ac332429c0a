"""Join an ncu --csv metric log of profiles/tools/forward_once.py with its operator sidecar (launch order) and fold it by
(kernel, operator, shape): measured time, DRAM bytes, L2 bytes, tensor-pipe activity per launch next to the ALGORITHMIC
FLOPs / bytes of that operator -> achieved TFLOP/s, achieved GB/s, traffic / algorithmic ratio.
  python profiles/tools/roofline_merge.py raw.csv oplog.json out.csv [peaks.json]"""
import csv
import io
import json
import re
import sys


def read_ncu(path):
    txt = open(path, errors="replace").read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    per = {}
    order = []
    for r in rows:
        if not r.get("ID", "").isdigit():
            continue
        i = int(r["ID"])
        if i not in per:
            per[i] = dict(name=r["Kernel Name"])
            order.append(i)
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1e-3)
        if "bytes" in m:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
        per[i][m] = v
    return [per[i] for i in order]


def main():
    raw, oplog, out = sys.argv[1:4]
    peaks = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else dict(hbm_gbs=6576.4, bf16_tflops=1680.1, bf16_tflops_sustained=1433.0)
    kernels = [k for k in read_ncu(raw) if "vx::" in k["name"] or k["name"].startswith("vx")]
    ops = json.load(open(oplog))
    if len(kernels) != len(ops):
        print(f"WARNING: {len(kernels)} vx kernels in the ncu log vs {len(ops)} logged operator launches; joining the common prefix")
    agg = {}
    for kr, op in zip(kernels, ops):
        kname = re.sub(r"\(.*", "", kr["name"]).replace("vx::", "").replace("void ", "")
        key = (op["part"], kname, op["op"], op["shape"])
        a = agg.setdefault(key, dict(n=0, us=0.0, dr=0.0, dw=0.0, lts=0.0, tens=0.0, flop=0.0, bytes=0.0))
        a["n"] += 1
        a["us"] += kr.get("gpu__time_duration.sum", 0.0)
        a["dr"] += kr.get("dram__bytes_read.sum", 0.0)
        a["dw"] += kr.get("dram__bytes_write.sum", 0.0)
        a["lts"] += kr.get("lts__t_bytes.sum", 0.0)
        a["tens"] += kr.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["flop"] += op["flop"] or 0.0
        a["bytes"] += op["bytes"] or 0.0
    rows = []
    for (part, kname, op, shape), a in agg.items():
        s = a["us"] * 1e-6
        rows.append(dict(part=part, kernel=kname, op=op, shape=shape, launches=a["n"], us_total=round(a["us"], 1),
                         us_per_launch=round(a["us"] / a["n"], 2),
                         algorithmic_gflop_per_launch=round(a["flop"] / a["n"] / 1e9, 3),
                         achieved_tflops=round(a["flop"] / s / 1e12, 1) if s and a["flop"] else "",
                         frac_of_sustained_tensor_peak=round(a["flop"] / s / 1e12 / peaks["bf16_tflops_sustained"], 3) if s and a["flop"] else "",
                         tensor_pipe_active_pct=round(a["tens"] / a["n"], 1),
                         algorithmic_mb_per_launch=round(a["bytes"] / a["n"] / 1e6, 2),
                         dram_mb_per_launch=round((a["dr"] + a["dw"]) / a["n"] / 1e6, 2),
                         dram_over_algorithmic=round((a["dr"] + a["dw"]) / a["bytes"], 2) if a["bytes"] else "",
                         l2_mb_per_launch=round(a["lts"] / a["n"] / 1e6, 2),
                         achieved_algorithmic_gbs=round(a["bytes"] / s / 1e9, 0) if s and a["bytes"] else "",
                         frac_of_hbm_peak=round(a["bytes"] / s / 1e9 / peaks["hbm_gbs"], 3) if s and a["bytes"] else ""))
    rows.sort(key=lambda r: -r["us_total"])
    with open(out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)
    tot = {}
    for r in rows:
        tot[r["part"]] = tot.get(r["part"], 0.0) + r["us_total"]
    print(f"{len(rows)} (kernel, op, shape) rows -> {out}; serialised time per part (us): {tot}")
    for r in rows[:12]:
        print(r["part"], r["kernel"][:34], r["op"], r["shape"], "n=%d" % r["launches"], "%.1f us" % r["us_per_launch"],
              "TF/s", r["achieved_tflops"], "tensor%", r["tensor_pipe_active_pct"], "dram/alg", r["dram_over_algorithmic"])


if __name__ == "__main__":
    main()
