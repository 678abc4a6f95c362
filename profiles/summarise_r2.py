"""Regenerate the round-2 summaries in profiles/ from the ncu reports / CSVs / JSON lines in
gpurun_out/ (scratch, not committed):

    python profiles/summarise_r2.py

Also writes profiles/ncu_traffic.json, the table bench.py reads its `roofline.traffic` from: DRAM
bytes per launch of the dominant kernel from the `ncu --set full` capture of the CURRENT kernel
sources (the file carries their hash; bench.py reports null when the sources have changed)."""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
sys.path.insert(0, ROOT)

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
]

# report -> what it is
REPORTS = [
    ("prof_r2_bx_f0", "v1: one CTA per 2048-pixel block, CTA-wide tile updated with shared-memory "
                      "atomicAdd (CAS loop), records loaded directly (pass 1 / pass 2)"),
    ("prof_r2b_bx_f0", "v2: the same with the records staged by cp.async.bulk + mbarrier "
                       "(1024-pixel blocks, 512-record stages): DRAM latency gone, pass 1 bound by "
                       "the shared-memory pipe (ATOMS.CAST.SPIN)"),
    ("prof_r2b_bx_f1", "v2 fused kernel"),
    ("prof_r2d_bx_f0", "v3: warp-private 256-pixel tiles, plain read-modify-write, direct REDs "
                       "(pass 1 / pass 2)"),
    ("prof_r2d_bx_f1", "v3 fused kernel"),
    ("prof_r2f_bx_f1", "v3 + persistent warps with a ticket counter (slower: 0.88 ms, second "
                       "sweep misses the L2 more often; dropped)"),
    ("prof_r2final_fused", "SHIPPED: warp-private 128-pixel tiles, static grid, three-buffer "
                           "software pipeline -- the capture bench.py's roofline.traffic points at"),
]


CAPTURE = None


def ncu_rows(rep):
    path = os.path.join(SRC, rep + ".ncu-rep")
    if not os.path.exists(path):
        return None, None
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[2:]


def ncu_summary():
    lines = ["# ncu --set full --clock-control none --import-source on, C4 shard (128 det x 2.16e6 "
             "samples, nside 2048)", ""]
    traffic = None
    for rep, what in REPORTS:
        hdr, rows = ncu_rows(rep)
        if hdr is None:
            continue
        stall = [h for h in hdr
                 if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", h)]
        lines.append(f"## {rep}: {what}")
        seen = set()
        for row in rows:
            name = row[hdr.index("Kernel Name")][:64]
            if name in seen:
                continue
            seen.add(name)
            lines.append("### " + name)
            for w in WANT:
                if w in hdr:
                    lines.append(f"{w:92s} {row[hdr.index(w)]}")
            for w in stall:
                try:
                    v = float(row[hdr.index(w)])
                except ValueError:
                    continue
                if v >= 0.3:
                    short = w.replace("smsp__average_warps_issue_stalled_", "stall ").replace(
                        "_per_issue_active.ratio", "")
                    lines.append(f"{short:92s} {v:.2f}")
            if rep == "prof_r2final_fused":
                unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                hu = ncu_units(rep)
                rd = float(row[hdr.index("dram__bytes_read.sum")]) * unit[hu["dram__bytes_read.sum"]]
                wr = float(row[hdr.index("dram__bytes_write.sum")]) * unit[hu["dram__bytes_write.sum"]]
                traffic = rd + wr
                pick = {"time_us": "gpu__time_duration.sum",
                        "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                        "l1_lsu_wavefronts_pct_of_peak":
                            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                        "of_which_shared_memory_pct":
                            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                        "l2_throughput_pct_of_peak":
                            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                        "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
                        "issue_slots_busy_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
                        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
                        "registers_per_thread": "launch__registers_per_thread"}
                global CAPTURE
                CAPTURE = {k: round(float(row[hdr.index(v)].replace(",", "")), 2)
                           for k, v in pick.items() if v in hdr}
                CAPTURE["bound_by"] = "L1/LSU data pipe (shared-memory tile updates + gathers)"
        lines.append("")
    open(os.path.join(OUT, "r2_ncu_blocked.txt"), "w").write("\n".join(lines) + "\n")
    return traffic


def ncu_units(rep):
    path = os.path.join(SRC, rep + ".ncu-rep")
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return dict(zip(rows[0], rows[1]))


def launch_list():
    path = os.path.join(SRC, "launches_r2j.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    i_n, i_v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[i_n].split("(")[0].replace("void ", "").replace("<unnamed>::", ""),
            float(r[i_v].replace(",", "")) / 1000.0) for r in rows[1:]]
    # the last three PCG iterations of the run: from the third k_pcg_update from the end
    upd = [i for i, (n, _) in enumerate(seq) if n.startswith("k_pcg_update")]
    out = ["kernel,launches_in_3_iterations,mean_us,total_us,share_of_step"]
    if len(upd) >= 4:
        window = seq[upd[-4]:upd[-1]]
        agg = collections.OrderedDict()
        for n, v in window:
            agg.setdefault(n[-70:], []).append(v)
        tot = sum(v for _, v in window)
        for n, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
            out.append(f"\"{n}\",{len(v)},{sum(v) / len(v):.1f},{sum(v):.1f},{sum(v) / tot:.3f}")
        out.append(f"\"TOTAL (3 iterations, cold-cache serialised ncu times)\",,,{tot:.1f},1.000")
    open(os.path.join(OUT, "r2_launches_step.csv"), "w").write("\n".join(out) + "\n")
    # the whole capture (set-up of the C4 shard + RHS + the first iterations; ncu -c 400): what
    # each set-up kernel costs once per solve
    agg = collections.OrderedDict()
    for n, v in seq:
        agg.setdefault(n[-70:], []).append(v)
    out = ["kernel,launches,mean_us,total_us"]
    for n, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        if sum(v) >= 100.0:
            out.append(f"\"{n}\",{len(v)},{sum(v) / len(v):.1f},{sum(v):.1f}")
    open(os.path.join(OUT, "r2_launches_setup.csv"), "w").write("\n".join(out) + "\n")


def copy_lines():
    for src, dst in (("bench_r2final.json", "r2_bench_n1.json"),
                     ("r2n2_auto.json", "r2_bench_n2_auto.json"),
                     ("r2n2_serial.json", "r2_bench_n2_serial.json"),
                     ("r2n2_nccl.json", "r2_bench_n2_nccl.json"),
                     ("r2n8_auto.json", "r2_bench_n8_auto.json"),
                     ("r2n8_serial.json", "r2_bench_n8_serial.json"),
                     ("r2n8_p2p.json", "r2_bench_n8_p2p.json"),
                     ("parity_fullsize_c4.json", "r2_parity_fullsize_c4.json"),
                     ("parity_fullsize_c3.json", "r2_parity_fullsize_c3.json"),
                     ("parity_fullsize_c5.json", "r2_parity_fullsize_c5.json"),
                     ("prior_timing_r2.txt", "r2_prior_timing.txt")):
        p = os.path.join(SRC, src)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(OUT, dst))
    parts = []
    for name in ("exp_r2_summary.txt", "r2a_summary.txt", "r2b_summary.txt", "r2c_summary.txt",
                 "r2d_summary.txt", "r2e_summary.txt", "r2f_summary.txt", "r2g_summary.txt",
                 "r2n2_summary.txt", "r2n2b_summary.txt", "r2n2c_summary.txt", "r2n8_summary.txt"):
        p = os.path.join(SRC, name)
        if os.path.exists(p):
            parts.append(f"## {name}\n" + open(p).read().strip() + "\n")
    open(os.path.join(OUT, "r2_variants.txt"), "w").write("\n".join(parts))


if __name__ == "__main__":
    traffic = ncu_summary()
    launch_list()
    copy_lines()
    if traffic is not None:
        import bench

        json.dump({"source_sha16": bench.source_hash(),
                   "sources": ["toast_b200/csrc/tb_blocked.cu", "toast_b200/csrc/tb_device.cuh",
                               "toast_b200/csrc/tb_obs.cuh"],
                   "kernels": {"k_bx<2>|c4|n1": {
                       "dram_bytes": traffic,
                       "ncu": CAPTURE,
                       "report": "profiles/r2_ncu_blocked.txt (prof_r2final_fused: ncu --set full "
                                 "--clock-control none, C4 shard, one B200)"}}},
                  open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)
        print("traffic", traffic, "hash", bench.source_hash())
