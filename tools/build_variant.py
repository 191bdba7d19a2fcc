"""Build a tuning variant of the library: python tools/build_variant.py name file.cu -DX=1 ...  ->
numgrids_b200/lib/variants/libng_<name>.so (select it with NUMGRIDS_B200_LIB=...)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numgrids_b200 import build as B  # noqa: E402

name, src, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
vdir = os.path.join(B.LIBDIR, "variants")
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(B.OBJDIR, f"{src[:-3]}_{name}.o")
r = subprocess.run([B.NVCC, *B.FLAGS, *defs, "-Xptxas=-v", "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
objs = [obj if s == src else os.path.join(B.OBJDIR, s.replace(".cu", ".o")) for s in B.SOURCES]
objs += [os.path.join(B.OBJDIR, o) for o in B.EXTRA_OBJECTS]
out = os.path.join(vdir, f"libng_{name}.so")
subprocess.run([B.NVCC, "-shared", "-o", out, *objs, "-cudart", "static"], check=True)
spill = [l for l in r.stderr.splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
print(out, f"({len(spill)} kernels spill)")
