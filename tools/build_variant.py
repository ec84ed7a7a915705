"""A/B build of libcoati_gpu.so with extra -D flags: python tools/build_variant.py NAME -DX=1 ... -> tools/gpu/ab_NAME.so
(run with COATI_GPU_LIB=$PWD/tools/gpu/ab_NAME.so; the .so travels to the GPU box, it is not committed)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coati_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu", "ab_%s.so" % name)
cus = sorted(os.path.join(B.CSRC, f) for f in os.listdir(B.CSRC) if f.endswith(".cu"))
ccs = sorted(os.path.join(B.CSRC, "host", f) for f in os.listdir(os.path.join(B.CSRC, "host")) if f.endswith(".cc") and f != "cli_main.cc")
subprocess.check_call([B.NVCC] + B.NVCC_FLAGS + defs + ["-o", out] + cus + ccs)
print(out)
