"""Per-opcode aggregation of the source page of an `ncu --set full --import-source on` capture
(`ncu -i REPORT --page source --csv`): where the warp instructions, the warp-state samples, the
shared-memory wavefronts and the global sectors of a kernel go.

    python profiles/ncu_source_mix.py gpurun_out/prof_r2final_fused.ncu-rep > profiles/r2_ncu_blocked_source.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernel, hdr, data = rows[0][1], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]].replace(",", ""))
        except (ValueError, IndexError):
            return 0.0

    tot = {k: sum(f(r, k) for r in data) for k in
           ("# Samples", "Instructions Executed", "L1 Wavefronts Shared",
            "L1 Wavefronts Shared Ideal", "L2 Theoretical Sectors Global",
            "L2 Theoretical Sectors Global Ideal", "L1 Tag Requests Global")}
    print(f"# {kernel}")
    print(f"# source: {report} (ncu --set full --clock-control none --import-source on, C4 shard)")
    print("warp instructions executed      %14.0f" % tot["Instructions Executed"])
    print("warp-state samples              %14.0f" % tot["# Samples"])
    print("shared-memory wavefronts        %14.0f  (ideal %.0f: %.0f %% excess = bank conflicts)" % (
        tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Ideal"],
        100.0 * (tot["L1 Wavefronts Shared"] / max(tot["L1 Wavefronts Shared Ideal"], 1) - 1)))
    print("global sectors (L2, theoretical) %13.0f  (ideal %.0f)" % (
        tot["L2 Theoretical Sectors Global"], tot["L2 Theoretical Sectors Global Ideal"]))
    print("global L1 tag requests          %14.0f" % tot["L1 Tag Requests Global"])
    byop = collections.defaultdict(lambda: [0.0] * 4)
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[ix["Source"]])
        b = byop[m.group(2) if m else "?"]
        b[0] += f(r, "Instructions Executed")
        b[1] += f(r, "# Samples")
        b[2] += f(r, "L1 Wavefronts Shared")
        b[3] += f(r, "L2 Theoretical Sectors Global")
    print("\n%-8s %13s %7s %9s %8s %13s %13s" % ("opcode", "warp inst", "% inst", "samples",
                                                 "% smpl", "sh wavefronts", "gl sectors"))
    for op, b in sorted(byop.items(), key=lambda x: -x[1][1])[:24]:
        print("%-8s %13.0f %6.1f%% %9.0f %7.1f%% %13.0f %13.0f" % (
            op, b[0], 100 * b[0] / tot["Instructions Executed"], b[1],
            100 * b[1] / tot["# Samples"], b[2], b[3]))
    st = {k: sum(f(r, k) for r in data) for k in hdr
          if k.startswith("stall_") and "(Not Issued)" not in k}
    s = sum(st.values())
    print("\nwarp-state samples by reason: " + ", ".join(
        "%s %.1f %%" % (k[6:], 100 * v / s) for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
