"""Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the stem kernels from an `ncu --set full` capture
of tools/prof_step.py, written as profiles/traffic.json (read by bench.py's roofline leg).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep [profiles/traffic.json]
"""
import csv, io, json, re, statistics, subprocess, sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                      "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[ki]
    rd = float(r[ir]) * scale[units[ir]]
    wr = float(r[iw]) * scale[units[iw]]
    key = None
    m = re.search(r"fast_fwd_kernel<\(?int\)?(\d)", name) or re.search(r"fast_fwd_kernel<(\d)", name)
    if m:
        key = f"conv{m.group(1)}.qconv_fwd_kernel"
    m = re.search(r"fast_bwd_pre_kernel<\(?int\)?(\d)", name) or re.search(r"fast_bwd_pre_kernel<(\d)", name)
    if m:
        key = f"conv{m.group(1)}.qconv_bwd_pre_kernel"
    if "stem_train_fwd_kernel" in name:
        key = "stem.qconv_fwd_kernel"
    if "fast_bwd_gy2_kernel" in name or "fast_bwd_gy_kernel" in name or "fast_bwd_gy3_kernel" in name:
        key = "gy"
    if key:
        acc.setdefault(key, []).append((rd + wr, rd, wr, float(r[it])))
res = {}
if "gy" in acc:  # the two layers share one instantiation: the conv1 launches read twice the bytes of the conv2 launches
    g = sorted(acc.pop("gy"))
    half = len(g) // 2
    acc["conv2.qconv_bwd_post_kernel"], acc["conv1.qconv_bwd_post_kernel"] = g[:half] or g, g[half:] or g
for k, v in acc.items():
    res[k] = round(statistics.median(t[0] for t in v))
    res[k + ".detail"] = {"launches": len(v), "dram_read": round(statistics.median(t[1] for t in v)),
                          "dram_write": round(statistics.median(t[2] for t in v)), "ncu_time_ns_or_us": statistics.median(t[3] for t in v)}
res["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, batch 16; writes still "
                "resident in the 126 MB L2 when the kernel ends are not counted by dram__bytes_write, so kernels that mostly "
                "WRITE (forward y, grad_x) show less traffic than their algorithmic bytes")
res["_source"] = rep
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
