import json, sys
d = json.load(open(sys.argv[1]))
print("value %.0f Msmp/s  step %.2f ms | e2e %.0f (%.2f ms) | dec %.0f (%.2f ms) e2e dec %.0f" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["decode"]["value"],
    d["decode"]["ms_per_step"], d["e2e"]["decode_value"]))
print("enc", {k.split()[-1][2:]: round(v, 2) for k, v in d["kernels_ms"].items() if v > 0.05})
print("dec", {k.split()[-1][2:]: round(v, 2) for k, v in d["decode"]["kernels_ms"].items()})
print(d["bit_exact"], d.get("cpu_baseline", {}).get("value"), d["clocks"], "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
