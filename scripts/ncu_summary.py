"""Condense `ncu --set full` reports (gpurun_out/*.ncu-rep) into the text summary kept under profiles/.
usage: python scripts/ncu_summary.py title=report.ncu-rep [...]  > profiles/rNN_ncu_full_summary.txt"""
import csv, io, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "smsp__inst_executed.sum",
]
STALL = "smsp__average_warps_issue_stalled_"

for arg in sys.argv[1:]:
    title, path = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("==== %s (%s)" % (title, path.split("/")[-1]))
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if name in seen:
            continue          # one launch per kernel is enough; repeats agree within 2 %
        seen.add(name)
        print("== %s" % name)
        for k in KEEP:
            if k in idx:
                print("  %-66s %s %s" % (k, r[idx[k]], units[idx[k]]))
        stalls = []
        for h, i in idx.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warps stalled per issue-active cycle, top reasons: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))
    print()
