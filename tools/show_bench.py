"""one line per bench JSON file: python tools/show_bench.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([x for x in open(f).read().splitlines() if x.startswith("{")][-1])
        r = d.get("roofline") or {}
        print("%-44s %-9s value %9.0f e2e %9.0f %s  %7.2f ms/step  frac %.3f  equiv %.0f TF  clocks %s" % (
            f.split("/")[-1], d.get("impl", "ours"), d["value"], d["e2e"]["value"], d["unit"], d["ms_per_step"], r.get("frac", 0),
            r.get("step_reference_equiv_tflops", r.get("achieved", 0)), (d.get("clocks") or {}).get("reasons")))
    except Exception as e:
        print(f, "unreadable:", e)
