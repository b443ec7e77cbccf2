"""Build the CUDA shared library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libselfmask_b200.so")
SOURCES = ["smk_model.cu", "smk_simt.cu", "smk_eval.cu", "smk_gemm_tc.cu", "smk_gemm_ln.cu", "smk_attn_tc.cu", "smk_attn_tc_multi.cu", "smk_attn_small.cu", "smk_attn_fa.cu", "smk_dec_attn.cu", "smk_xattn_tc.cu", "smk_mask_mma.cu", "smk_finalize.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]
if os.environ.get("SMK_BUILD_TUNE") == "1":      # tuning build: GEMM wait-cycle trace + stage-isolation switches (scripts/gemm_trace.py)
    NVCC_FLAGS.append("-DSMK_GEMM_TUNE=1")


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def source_hash():
    """Content hash of every source the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "selfmask_b200.h")]
    for d in deps:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    stamp = LIB + ".srchash"
    return not (os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == source_hash())


def build(force=False, verbose=False):
    """Compile every kernel for sm_100a into libselfmask_b200.so (parallel per-file, then link)."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    missing = [s for s in SOURCES if not os.path.exists(os.path.join(CSRC, s))]
    if missing:
        raise RuntimeError(f"listed CUDA sources are missing: {missing}")
    srcs = list(SOURCES)
    procs = []
    for s in srcs:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for s, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        objs.append(obj)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(LIB + ".srchash", "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
