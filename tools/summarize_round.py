"""python tools/summarize_round.py profiles/r2_bench_n1.json [profiles/r2_bench_n2.json ...] : the numbers DESIGN.md 8 quotes"""
import json, sys
base = None
for f in sys.argv[1:]:
    l = json.loads(open(f).read().strip().splitlines()[-1])
    n = l["n_gpus"]
    r = l["roofline"]
    ph = r.get("phase_ms_per_step") or r.get("phase_ms_per_step_rank0") or {}
    if n == 1:
        base = l["value"]
    eff = " eff %.3f" % (l["value"] / (n * base)) if base else ""
    print("N=%d %s: %.1f ms/step  %.4g k-mers/s%s | e2e %.1f ms %.4g/s | parity %s | launches %d" % (
        n, l["config"]["name"], l["ms_per_step"], l["value"], eff, l["e2e"]["ms_per_step"], l["e2e"]["value"], l["parity"]["match"], l["gpu_launches"]))
    print("    phases", {k[3:]: round(v, 1) for k, v in ph.items()}, "sum %.1f" % sum(ph.values()))
    if "host_wall_ms_per_step_rank0" in r:
        print("    host wall", {k: round(v, 1) for k, v in r["host_wall_ms_per_step_rank0"].items()})
    if n == 1:
        print("    roofline", r["kernel"], "achieved %.0f GB/s frac %.3f traffic %s | pipeline %.0f GB/s frac %.3f" % (
            r["achieved"], r["frac"], r.get("traffic"), r["pipeline"]["achieved"], r["pipeline"]["frac"]))
        print("    e2e", {k: v for k, v in l["e2e"].items() if k not in ("value", "unit")})
        print("    cpu_baseline", l.get("cpu_baseline"))
        if "secondary" in l:
            s = l["secondary"]
            print("    secondary %s: %.2f ms %.4g/s e2e %.2f ms %.4g/s parity %s" % (s["config"]["name"], s["ms_per_step"], s["value"], s["e2e"]["ms_per_step"], s["e2e"]["value"], s["parity"]["match"]))
        if "unitig" in l:
            u = l["unitig"]; print("    unitig %.3f s, cpu %s" % (u["seconds"], u.get("cpu_baseline")))
        if "cli" in l:
            print("    cli", l["cli"])
