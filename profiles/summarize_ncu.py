"""Turns an ncu report (gpurun_out/*.ncu-rep, scratch) into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/aec_prof_r1.ncu-rep profiles/r1a_aec_kernel

writes <out>.metrics.txt (selected raw-page metrics per captured launch) and <out>.hotlines.txt (source lines ranked by
warp-stall samples and by executed instructions; needs -lineinfo, which the Makefile passes)."""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(out.splitlines()))


def main(rep, out):
    rows = ncu_csv(rep, "--page", "raw")
    H, units = rows[0], rows[1]
    with open(out + ".metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none; source report: {rep}\n")
        for r in rows[2:]:
            d = dict(zip(H, r))
            f.write(f"\nkernel {d.get('Kernel Name', '')[:60]} grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for w in WANT:
                if w in d:
                    f.write(f"  {w:70s} {d[w]:>18s} {units[H.index(w)]}\n")
            try:
                rd = float(d["dram__bytes_read.sum"].replace(",", ""))
                wr = float(d["dram__bytes_write.sum"].replace(",", ""))
                f.write(f"  {'dram traffic (read+write) as reported':70s} {rd + wr:>18.4f} {units[H.index('dram__bytes_read.sum')]}\n")
            except (KeyError, ValueError):
                pass
    rows = ncu_csv(rep, "--page", "source", "--print-source", "sass,cuda")
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    H = rows[hdr]
    idx = {h: i for i, h in enumerate(H)}
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(H) or r[2] != "-":
            continue
        try:
            ln, ins, smp = int(r[0]), int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]])
        except ValueError:
            continue
        a = agg.setdefault(ln, [0, 0, r[1][:110], collections.Counter()])
        a[0] += ins
        a[1] += smp
        for h in H:
            if h.startswith("stall_") and "Not Issued" not in h and r[idx[h]].isdigit():
                a[3][h] += int(r[idx[h]])
    ti, ts = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
    with open(out + ".hotlines.txt", "w") as f:
        f.write(f"# per source line (all captured launches summed); samples={ts} warp-instructions={ti}\n== by stall samples\n")
        for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            top = ", ".join(f"{k}:{v}" for k, v in a[3].most_common(2))
            f.write(f"smp {a[1] / ts:6.1%} inst {a[0] / ti:6.1%} L{ln:4d} {a[2]}   [{top}]\n")
        f.write("== by executed instructions\n")
        for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
            f.write(f"inst {a[0] / ti:6.1%} smp {a[1] / ts:6.1%} L{ln:4d} {a[2]}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
