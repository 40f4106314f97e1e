"""In-tree build of libppasr_b200.so (hand-written sm_100a CUDA + the C-ABI).

`python -m ppasr_b200.build` compiles every .cu under ppasr_b200/csrc with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
and links ppasr_b200/lib/libppasr_b200.so. nvcc cross-compiles without a GPU; the .so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
OUT_DIR = os.path.join(ROOT, "lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB_PATH = os.path.join(OUT_DIR, "libppasr_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(os.path.dirname(ROOT), "include"),
    "-I", CSRC,
]


def _newer(src, dst, deps):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in [src] + deps)


def build(verbose=False, force=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inl"))]
    inc = os.path.join(os.path.dirname(ROOT), "include")
    if os.path.isdir(inc):
        headers += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h")]
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            jobs.append([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    failed = False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd[-4:]) + "\n" + r.stdout + r.stderr + "\n")
            if r.returncode != 0:
                failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    if jobs or not os.path.exists(LIB_PATH) or force:
        cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
