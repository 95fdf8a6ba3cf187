"""Kernel shares of the step from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file ...).

    python tools/launch_shares.py gpurun_out/launches_bench.csv > profiles/<name>.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's `kernels[*].share`, not absolutes."""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ui, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name")
agg = collections.OrderedDict()
seq = []
for r in rows[hdr + 1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("qw::", "")
    name = re.sub(r"^(st|lm|gen|wc)::", "", name)  # kernels of the library's nested namespaces
    agg.setdefault(name, []).append(t)
    seq.append((name, t))
own = {k: v for k, v in agg.items() if k.startswith(("fast_", "qconv_", "logmel_", "wcirc", "circuit_", "gen_", "grads_", "stem"))}
tot = sum(sum(v) for v in own.values())
print(f"# {sys.argv[1]}: {len(seq)} launches, {sum(len(v) for v in own.values())} from libqw_b200.so; times in us (ncu, cold cache, serialised)")
print(f"{'kernel':58s} {'launches':>8s} {'mean us':>9s} {'share of own time':>18s}")
for k, v in sorted(own.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:58s} {len(v):8d} {sum(v)/len(v):9.2f} {sum(v)/tot:18.3f}")
other = {k: v for k, v in agg.items() if k not in own}
print(f"# other (torch) kernels: {sum(len(v) for v in other.values())} launches, {sum(sum(v) for v in other.values()):.1f} us total")
