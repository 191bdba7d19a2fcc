"""Static SASS opcode histograms of the hot kernels (cuobjdump -sass on the built objects): the evidence for which
hardware paths a kernel uses (UTMALDG / UBLKCP = TMA and bulk copies, SYNCS = mbarrier, LDGSTS = cp.async, DMUL /
DADD / DFMA = FP64 pipe; no HMMA / DMMA / UTCMMA anywhere: the parity contract forbids reordered sums).

    python tools/sass_hist.py > profiles/r02b_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "numgrids_b200", "build")
HOT = [("fdlap3d.o", r"fdlap3d_ring2_kernel"), ("curvlap3d.o", r"curvlap3d_kernel"), ("fd.o", r"fd_march_kernel"),
       ("fd_row.o", r"fd_row_kernel"), ("dense.o", r"dense_fast_kernel"), ("fft2.o", r"fft2pass_ring_kernel"),
       ("fft2.o", r"fft2pass_kernel"), ("fft_tile.o", r"fft_tile_kernel"), ("fft.o", r"fft_diff_kernel"), ("fft_mixed.o", r"fft_mixed_kernel"), ("fd_tile.o", r"fd_tile_kernel"), ("slab.o", r"slab_"), ("quad.o", r"quad_partial_kernel"), ("transpose.o", r"slab_block_copy_kernel")]
KEY = ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "DMUL", "DADD", "DFMA", "HMMA", "DMMA", "UTCMMA", "LDS", "STS", "LDG",
       "STG", "SHFL", "ATOMS", "BAR", "MUFU", "LDL", "STL")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    cur, res = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            res[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            res[cur][m.group(1)] += 1
    return res


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


print("# static SASS opcode counts per kernel instance (sm_100a, the objects the shipped .so is linked from)")
for obj, pat in HOT:
    fns = {k: v for k, v in functions(obj).items() if re.search(pat, k)}
    # the instances a default run takes are among them; list at most 4 per family, largest first
    for name, cnt in sorted(fns.items(), key=lambda kv: -sum(kv[1].values()))[:4]:
        total = sum(cnt.values())
        full = demangle(name)
        short = full[:full.index(">(") + 1] if ">(" in full else full
        keys = "  ".join(f"{k}={cnt[k]}" for k in KEY if cnt[k])
        print(f"\n{obj}: {short}\n  {total} instructions ({total * 16 // 1024} KiB)  {keys}")
        print("  top: " + ", ".join(f"{k} {v}" for k, v in cnt.most_common(10)))
