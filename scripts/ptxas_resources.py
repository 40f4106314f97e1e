"""Static resource table of every kernel (no GPU needed): registers, stack, spills, static shared memory as reported by
`ptxas -v` for sm_100a.  python scripts/ptxas_resources.py > profiles/<round>_ptxas_resources.txt"""
import concurrent.futures
import glob
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ppasr_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-Xptxas", "-v"]


def one(src):
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", os.path.join(d, "o.o")], capture_output=True, text=True)
    rows = []
    for b in re.split(r"ptxas info\s+: Compiling entry function '", r.stderr)[1:]:
        name = subprocess.run(["c++filt", b.split("'")[0]], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        used = re.search(r"Used (\d+) registers", b)
        spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
        smem = re.search(r"(\d+) bytes smem", b)
        rows.append((os.path.basename(src), name[:64], used.group(1) if used else "?",
                     *(spill.groups() if spill else ("?",) * 3), smem.group(1) if smem else "0"))
    return rows


if __name__ == "__main__":
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    print("# nvcc " + " ".join(FLAGS[:8]) + " ... -Xptxas -v   (static; dynamic shared memory is set at launch)")
    print(f"{'file':22} {'kernel':64} {'regs':>5} {'stack':>6} {'spill_st':>8} {'spill_ld':>8} {'static_smem':>11}")
    with concurrent.futures.ThreadPoolExecutor(8) as ex:
        for rows in ex.map(one, srcs):
            for r in rows:
                print(f"{r[0]:22} {r[1]:64} {r[2]:>5} {r[3]:>6} {r[4]:>8} {r[5]:>8} {r[6]:>11}")
