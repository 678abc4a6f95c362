"""Static view of the shipped kernels (no GPU needed): registers / stack / shared memory from
`cuobjdump -res-usage` and the SASS instruction mix from `cuobjdump -sass`, for the kernels the
bench line and DESIGN.md talk about.

    python profiles/sass_static.py > profiles/r2_sass_static.txt

What to read off it: no kernel on the per-iteration path spills (STACK 0 / no STL, LDL); the
block-ordered passes contain MATCH.ANY + plain LDS/STS on the tile and RED.E.ADD.F64 only for the
amplitudes (no ATOMS: no shared-memory atomics); the NVLS reduction holds MULTIMEM.LD_REDUCE /
MULTIMEM.ST; fp64 products and sums are separate DMUL / DADD (-fmad=false: the reference's double
rounding) -- DFMA appears only inside the compiler's IEEE division / square-root sequences (with
their MUFU seeds) and in the explicit fma() calls of the double-double atan2 (tb_math.cuh).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "toast_b200", "libtoastb200.so")

WANT = [
    ("k_bx<2,1,1>", r"4k_bxILi2ELb1ELb1EE"),
    ("k_bx<0,1,1>", r"4k_bxILi0ELb1ELb1EE"),
    ("k_bx<1,1,1>", r"4k_bxILi1ELb1ELb1EE"),
    ("k_bx_prescale", r"13k_bx_prescale"),
    ("k_map_reduce_cov<8>", r"16k_map_reduce_covILi8EE"),
    ("k_map_reduce_cov<2>", r"16k_map_reduce_covILi2EE"),
    ("k_map_reduce_cov_mc", r"19k_map_reduce_cov_mc"),
    ("k_pcg_update", r"k_pcg_update"),
    ("k_pcg_direction", r"k_pcg_direction"),
    ("k_amp_dot", r"9k_amp_dot"),
    ("k_pointing_fused<NEST,noHWP>", r"16k_pointing_fusedILb1ELb0EE"),
    ("k_build_noise_weighted<3>", r"k_build_noise_weightedILi3E"),
    ("k_cov_accum<3>", r"k_cov_accumILi3E"),
    ("k_pixels_wcs", r"12k_pixels_wcs"),
    ("k_pb_chunk<1>", r"10k_pb_chunkILb1EE"),
    ("k_pb_seg<1>", r"8k_pb_segILb1EE"),
    ("k_prior_conv<0>", r"12k_prior_convILi0EE"),
]
GROUPS = collections.OrderedDict([
    ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"),
    ("RED", r"^RED"), ("ATOMG", r"^ATOMG|^ATOM\b|^ATOM\."), ("ATOMS", r"^ATOMS"),
    ("SHFL", r"^SHFL"), ("MATCH", r"^MATCH"), ("REDUX", r"^REDUX|^CREDUX"), ("BAR", r"^BAR"),
    ("LDL/STL", r"^LDL|^STL"), ("MULTIMEM", r"MULTIMEM"), ("DADD", r"^DADD"),
    ("DMUL", r"^DMUL"), ("DFMA", r"^DFMA"), ("MUFU", r"^MUFU"), ("UBLKCP", r"^UBLKCP"),
])


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in ln:
            usage[fn] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", ln))
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    mix, cur = {}, None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            mix[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur:
            op = m.group(1)
            mix[cur]["_total"] += 1
            for g, pat in GROUPS.items():
                if re.search(pat, op):
                    mix[cur][g] += 1
    print("# static SASS / resource view of toast_b200/libtoastb200.so (sm_100a), "
          "profiles/sass_static.py")
    hdr = ["kernel", "REG", "STACK", "SHARED", "SASS"] + list(GROUPS)
    print(" | ".join(hdr))
    for name, pat in WANT:
        fns = [f for f in mix if re.search(pat, f)]
        if not fns:
            print(name, "| (not found)")
            continue
        f = fns[0]
        u = usage.get(f, {})
        row = [name, u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"),
               str(mix[f]["_total"])] + [str(mix[f][g]) for g in GROUPS]
        print(" | ".join(row))
    spill = sorted((f for f, u in usage.items() if int(u.get("STACK", 0)) > 0),
                   key=lambda f: -int(usage[f]["STACK"]))
    print("\n# kernels with a stack frame (local arrays or spills):")
    for f in spill:
        short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_", "", f)[:90]
        print(f"{usage[f]['STACK']:>5} B  REG {usage[f]['REG']:>3}  {short}")


if __name__ == "__main__":
    sys.exit(main())
