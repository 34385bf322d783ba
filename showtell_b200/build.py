"""In-tree build of libshowtell_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m showtell_b200.build            # incremental
    python -m showtell_b200.build --force

The .so lands next to this file so it travels to the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, os.environ.get("SHOWTELL_B200_LIBNAME", "libshowtell_b200.so"))   # A/B builds: other name
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _digest(path):
    h = hashlib.sha1()
    for dep in [path] + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
            sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        with open(dep, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, force):
    # A/B builds: SHOWTELL_B200_VARIANT="file.cu:-DX=1 -DY=2" compiles that one source with extra flags into its own object
    var = os.environ.get("SHOWTELL_B200_VARIANT", "")
    extra = var.split(":", 1)[1].split() if var and os.path.basename(src) == var.split(":", 1)[0] else []
    tag = ("." + hashlib.sha1(" ".join(extra).encode()).hexdigest()[:8]) if extra else ""
    obj = os.path.join(OBJ, os.path.basename(src) + tag + ".o")
    stamp = obj + ".sha1"
    dig = _digest(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    r = subprocess.run([NVCC] + FLAGS + extra + ["-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [r[0] for r in res]
    if verbose:
        for r in res:
            if r[1]:
                print(r[2])
    if any(r[1] for r in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
