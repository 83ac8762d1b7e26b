"""Reduce an `ncu --page raw --csv` dump to profiles/traffic_r01.json: DRAM bytes per launch and
duration of every profiled kernel.  usage: python tools/ncu_traffic.py raw.csv channel_samples out.json"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(v) * scale
out, detail = {}, {}
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[col["Kernel Name"]])
    short = re.match(r"([A-Za-z0-9_]+)", name).group(1)
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    if short in out and out[short] >= rd + wr:
        continue            # template variants of one kernel (the exact-order fallbacks): keep the main one
    out[short] = rd + wr
    detail[short] = {"dram_read_bytes": rd, "dram_write_bytes": wr,
                     "gpu_time_ms": float(r[col["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[col["gpu__time_duration.sum"]]],
                     "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                     "registers": r[col["launch__registers_per_thread"]],
                     "issue_active_pct": r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]],
                     "warps_active_pct": r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]}
json.dump({"channel_samples": int(sys.argv[2]), "source": "ncu --set full --clock-control none, one launch per kernel",
           "dram_bytes_per_launch": out, "detail": detail}, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out))
