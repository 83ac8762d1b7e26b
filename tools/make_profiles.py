"""Turn the ncu dumps brought back in gpurun_out/ into the committed summaries under profiles/.
usage: python tools/make_profiles.py <launches.csv> <raw.csv> <src-prefix> <channel_samples> [tag]"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, raw, srcp, chsamp = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
TAG = sys.argv[5] if len(sys.argv) > 5 else "r02"
P = os.path.join(ROOT, "profiles")

# ---------------------------------------------------------------- launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
seq = collections.OrderedDict()
other = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[col["Kernel Name"]]); short = re.match(r"([A-Za-z0-9_:]+)", name).group(1)
    ns = float(r[col["Metric Value"]].replace(",", ""))
    (seq if short.startswith("k_") else other).setdefault(short, []).append(ns)
tot = sum(sum(v) for v in seq.values())
L = [f"# ncu launch list of the bench command, {TAG}", "",
     "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv python bench.py --steps 2 --warmup 3 --configs \"\" --no-cpu-baseline`",
     f"(raw CSV: `profiles/launches_{TAG}.csv`).  Times are cold-cache and serialised: compare shares, not absolutes.",
     "The list covers the whole run: warm-up, the profiled and the timed device-resident calls, and the pipelined host-API / PCM calls,",
     "which run every kernel once per chunk.", "",
     "| kernel | launches | total ms | share of our kernels |", "|---|---|---|---|"]
for k, v in sorted(seq.items(), key=lambda kv: -sum(kv[1])):
    L.append(f"| {k} | {len(v)} | {sum(v) / 1e6:.3f} | {100 * sum(v) / tot:.1f} % |")
L += ["", "Other launches in the process (PyTorch fills/copies used by bench.py to build inputs and compare outputs): " +
      ", ".join(f"{k} x{len(v)} ({sum(v) / 1e6:.2f} ms)" for k, v in other.items())]
enc = {k: v[:5] for k, v in seq.items() if k.startswith("k_enc") or k == "k_scan_u32"}
dec = {k: v[:3] for k, v in seq.items() if k.startswith("k_dec")}
def table(d, title):
    per = {k: sum(v) / len(v) * (2 if k == "k_scan_u32" else 1) for k, v in d.items()}
    t = sum(per.values())
    out = [f"### {title}", "", "| kernel | ms per whole-file call (ncu, cold, serialised) | share |", "|---|---|---|"]
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
        out.append(f"| {k} | {v / 1e6:.3f} | {100 * v / t:.1f} % |")
    out.append(f"| total | {t / 1e6:.3f} | |")
    return out
L += ["", "## Device-resident whole-file calls only", "",
      "The first launches of each kernel in the list (warm-up and profiled single-pass calls on the full 1-hour file; the later ones",
      "belong to chunked calls).  These shares are the ones to compare with `kernels_ms` / `roofline.kernel_share_of_step` in the bench line.", ""]
L += table(enc, "encode (C2: 317.5 M channel-samples)") + [""] + table(dec, "decode")
open(os.path.join(P, f"launches_{TAG}_summary.md"), "w").write("\n".join(L) + "\n")
subprocess.run(["cp", launches, os.path.join(P, f"launches_{TAG}.csv")], check=True)

# ---------------------------------------------------------------- full set
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), raw, str(chsamp), os.path.join(P, f"traffic_{TAG}.json")],
               check=True, stdout=subprocess.DEVNULL)
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]; col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/instr"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
def fmt(v):
    try:
        f = float(v.replace(",", ""))
        return f"{f:.3g}" if abs(f) < 1000 else f"{f:.0f}"
    except ValueError:
        return v
L = [f"# ncu --set full, {TAG} kernels at the bench size (C2: 1 h, 16-bit stereo, preset 2)", "",
     "Command: `ncu --set full --clock-control none --import-source on -k regex:\"k_enc_|k_dec_|k_scan\" -c 45 python tools/profile_run.py 3600 2`",
     "(one device-resident encode + decode of the 317.5 M channel-sample file; first launch of every kernel).  The .ncu-rep is 60+ MB and is",
     "not committed; `profiles/traffic_" + TAG + ".json` holds the DRAM bytes per launch that `bench.py` reports as `roofline.traffic`.", "",
     "| kernel | " + " | ".join(w[1] + (" (" + units[col[w[0]]] + ")" if units[col[w[0]]] and w[1] in ("time", "dram read", "dram write") else "") for w in want) + " |",
     "|---|" + "---|" * len(want)]
seen = set()
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[col["Kernel Name"]]); short = re.match(r"([A-Za-z0-9_:]+(<[^>]*>)?)", name).group(1)
    if short in seen:
        continue
    seen.add(short)
    L.append("| " + short + " | " + " | ".join(fmt(r[col[w[0]]]) for w in want) + " |")
notes = os.path.join(P, f"ncu_{TAG}_notes.md")
L += [""] + (open(notes).read().splitlines() if os.path.exists(notes) else []) + [""]
for k in ("k_enc_ltcorr_mma", "k_enc_pack_rice", "k_enc_ltlms", "k_enc_ricetrace", "k_dec_block", "k_enc_lagsums", "k_enc_parcor", "k_enc_ltfft"):
    f = f"{srcp}_{k}.csv"
    if not os.path.exists(f):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src_summary.py"), f, "6"], capture_output=True, text=True).stdout.splitlines()
    L += [f"### {k}: stall reasons (source page, warp samples)", "", "```"] + out[:9] + ["```", ""]
open(os.path.join(P, f"ncu_{TAG}_summary.md"), "w").write("\n".join(L) + "\n")
print("profiles written")
