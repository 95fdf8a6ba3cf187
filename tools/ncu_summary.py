"""Summarise an .ncu-rep: per kernel launch, key raw metrics + opcode mix + stall reasons + hottest SASS.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 12] [--kernel regex]
"""
import argparse
import collections
import csv
import io
import re
import subprocess

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--kernel", default=None)
    a = ap.parse_args()
    raw = run(["ncu", "-i", a.rep, "--page", "raw", "--csv", "--metrics", ",".join(RAW)])
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = rows[2:]
    src = run(["ncu", "-i", a.rep, "--page", "source", "--csv"])
    blocks = re.split(r'(?m)^"Kernel Name",', src)[1:]
    if len(blocks) == 2 * len(launches):
        blocks = blocks[::2]
    elif len(blocks) > len(launches):
        # kernels compiled with -lineinfo come out twice (identical SASS views), library kernels without it once:
        # drop exact consecutive duplicates until the counts agree
        dedup = []
        extra = len(blocks) - len(launches)
        for b in blocks:
            if extra and dedup and b == dedup[-1]:
                extra -= 1
                continue
            dedup.append(b)
        blocks = dedup
    for li, (row, blk) in enumerate(zip(launches, blocks)):
        name = row[hdr.index("Kernel Name")]
        if a.kernel and not re.search(a.kernel, name):
            continue
        print("=" * 110)
        print(f"[{li}] {name}")
        for m in RAW:
            if m in hdr:
                print(f"    {m:68s} {row[hdr.index(m)]:>16s} {units[hdr.index(m)]}")
        lines = blk.split("\n")
        rd = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        if not rd:
            continue
        h = rd[0]
        ci = {k: h.index(k) for k in h}
        ops = collections.Counter()
        stalls = collections.Counter()
        tot_inst = 0
        tot_samp = 0
        recs = []
        for r in rd[1:]:
            if len(r) < len(h):
                continue
            sass = r[ci["Source"]].strip()
            inst = int(r[ci["Instructions Executed"]] or 0)
            samp = int(r[ci["# Samples"]] or 0)
            op = sass.split()[0] if sass else "?"
            if op.startswith("@"):
                op = sass.split()[1]
            ops[op.split(".")[0]] += inst
            tot_inst += inst
            tot_samp += samp
            for k in h:
                if k.startswith("stall_") and "Not Issued" not in k:
                    stalls[k] += int(r[ci[k]] or 0)
            recs.append((samp, inst, sass))
        print(f"    -- warp instructions {tot_inst}, samples {tot_samp}")
        print("    -- opcode mix:", ", ".join(f"{k}:{v * 100 // max(1, tot_inst)}%" for k, v in ops.most_common(14)))
        print("    -- stalls:", ", ".join(f"{k[6:]}:{v * 100 // max(1, tot_samp)}%" for k, v in stalls.most_common(8)))
        print(f"    -- hottest SASS (samples, executed):")
        for samp, inst, sass in sorted(recs, reverse=True)[: a.top]:
            print(f"       {samp:7d} {inst:10d}  {sass[:90]}")


if __name__ == "__main__":
    main()
