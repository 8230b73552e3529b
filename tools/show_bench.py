"""Compact view of a bench.py JSON line: python tools/show_bench.py gpurun_out/x.json"""
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def show(tag, d):
    r = d["roofline"]
    print("%s %.4g k-mers/s  %.2f ms/step  e2e %.2f ms (%.3g/s)  pipeline frac %.3f  dom %s frac %.3f" % (
        tag, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"], r["pipeline"]["frac"], r["kernel"], r["frac"]))
    print("   phases", {k: round(v, 2) for k, v in r["phase_ms_per_step"].items()})
    print("   e2e", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k not in ("value", "unit")})
    print("   result", d["result"], "parity", d["parity"]["match"], d["parity"]["expected"] is not None)
show("PRIMARY  ", l)
if "secondary" in l: show("SECONDARY", l["secondary"])
if "unitig" in l: print("unitig", l["unitig"])
if "cpu_baseline" in l: print("cpu", l["cpu_baseline"]["value"], l["cpu_baseline"]["cores"])
print("clocks", l.get("clocks"))
