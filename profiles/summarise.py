"""Regenerate the text summaries in profiles/ from the ncu reports / CSVs in gpurun_out/.

    python profiles/summarise.py

(ncu is on the build container; the .ncu-rep files themselves are not committed.)"""
import collections
import csv
import io
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def ncu_summary(reports, out_name):
    lines = []
    for rep in reports:
        path = os.path.join(SRC, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        stall = [h for h in hdr
                 if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h)]
        lines.append("## " + rep)
        for row in rows[2:]:
            lines.append("### " + row[hdr.index("Kernel Name")][:70])
            for w in WANT + stall:
                if w not in hdr:
                    continue
                i = hdr.index(w)
                if w in stall:
                    try:
                        if float(row[i]) < 0.3:
                            continue
                    except ValueError:
                        pass
                lines.append("%-100s %s %s" % (w, row[i], units[i]))
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines) + "\n")


def launch_summary(csv_name, out_name, pass1_pattern, n_iter=3):
    lines = [ln for ln in open(os.path.join(SRC, csv_name)) if not ln.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]

    def us(r):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        return v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)

    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if re.search(pass1_pattern, n)]
    # launches of pass 1: prologue, warm-up ..., then the timed iterations; take iterations
    # 2..4 of the pipelined loop (update .. dot), i.e. from the update preceding idx[2]
    start, end = idx[2], idx[2 + n_iter]
    # move both bounds back to the k_pcg_update that opens the step
    def back(i):
        while i > 0 and "k_pcg_update" not in names[i]:
            i -= 1
        return i
    start, end = back(start), back(end)
    agg = collections.OrderedDict()
    for r in rows[start:end]:
        n = re.sub(r"\(.*", "", r["Kernel Name"])[:72]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us(r)
    tot = sum(v[1] for v in agg.values())
    n_iter = max(1, sum(1 for r in rows[start:end] if re.search(pass1_pattern, r["Kernel Name"])))
    out = ["# ncu launch list of %d timed PCG iterations (1 x B200, default workload)" % n_iter,
           "# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv "
           "python bench.py --steps 3 --warmup 1 --no-cpu-baseline",
           "# per-launch times under ncu are cold-cache and serialised: compare SHARES",
           "kernel,launches,total_us,share_pct,us_per_iteration"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s,%d,%.1f,%.1f,%.1f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / n_iter))
    out.append("TOTAL,,%.1f,100.0,%.1f" % (tot, tot / n_iter))
    open(os.path.join(OUT, out_name), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    import sys
    if "--r1" in sys.argv:  # first two sessions (per-sample kernels)
        ncu_summary(["prof_r1_passes", "prof_r1_tma", "prof_r1_compact", "prof_r1_pair"],
                    "r1_ncu_passes.txt")
        launch_summary("launches_r1.csv", "r1_launches_step.csv", r"k_lhs_pair<\(?(bool\))?0>")
    else:  # session 3: crossing-list kernels (the shipped path)
        # prof_s3_x: k_bin_xs + time-ordered k_lhs_x<1>; prof_s4_x: k_bin_xs + k_proj_xs (shipped)
        ncu_summary(["prof_s3_x", "prof_s4_x"], "r1_ncu_crossings.txt")
        launch_summary("launches_s4.csv", "r1_launches_step_crossings.csv", r"k_bin_xs")
