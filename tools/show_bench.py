import json, signal, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # quiet under `| head`
d = json.load(open(sys.argv[1]))
print("value %.4g %s  ms/step %.5f  step_frac %.4f  e2e %.4g (%.4f ms)" % (d["value"], d["unit"], d["ms_per_step"], d["step_roofline_frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "kernel_ms", "kernel_share_of_step", "traffic")})
for k, v in d["kernels"].items():
    print("   %-36s %8.2f us  share %.3f  frac %s" % (k, v["ms"] * 1e3, v["share"], v["frac"]))
print("calls", d["calls_ms"])
print("clocks", d["clocks"])
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("vectorised_oracle"))
if d.get("encoder_fwd"): print("encoder", d["encoder_fwd"])
