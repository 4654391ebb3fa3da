"""Build a variant of libmarl_b200.so with extra -D flags for kernel A/B runs:  ab_build.py NAME -DMARL_FETCH_DEPTH=2 ...
The library lands in marl_b200/lib/ab/libmarl_NAME.so (git-ignored, travels to the GPU box); select it with
MARL_B200_LIB=marl_b200/lib/ab/libmarl_NAME.so."""
import os, subprocess, sys, concurrent.futures as cf
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from marl_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(B.OUT_DIR, "ab", name)
os.makedirs(out, exist_ok=True)

def one(src):
    obj = os.path.join(out, os.path.basename(src)[:-3] + ".o")
    r = subprocess.run([B.NVCC, *B.FLAGS, *defs, "-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    return obj

with cf.ThreadPoolExecutor(max_workers=8) as ex:
    objs = list(ex.map(one, B.sources()))
lib = os.path.join(B.OUT_DIR, "ab", f"libmarl_{name}.so")
subprocess.run([B.NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
for o in objs:
    os.remove(o)
print(lib)
