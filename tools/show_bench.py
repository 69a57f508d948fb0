import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline") or {}
print("value %.1f samples/s  %.2f ms/step  e2e %.1f  gemm %.1f TF/s (%.2f ms/step)  step-frac %.3f  launches %s  loss %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("achieved", 0), r.get("gemm_ms_per_step", 0), r.get("step_executed_frac_of_peak", r.get("step_frac_of_peak", 0)),
    d.get("gpu_launches"), d.get("loss_first_last")))
