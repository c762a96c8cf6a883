"""In-tree build of the C-ABI CUDA library (sm_100a only).  No JIT cache: the .so sits
next to the package so that it travels with the source tree."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libalad_b200.so")
SOURCES = ["cabi.cu", "pack.cu", "mrsw_fwd.cu", "mrsw_fwd_tf32.cu", "fused.cu", "mrsw_bwd.cu", "losses.cu", "train_step.cu", "distill.cu", "misc_sim.cu", "scan_pool.cu", "rank.cu", "retrieval.cu", "pairs.cu", "peer.cu", "h2d.cu"]
NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--threads", "4",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build aladin_b200/libalad_b200.so")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "alad_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into aladin_b200/libalad_b200.so."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    if os.environ.get("ALAD_NVCC_EXTRA"):               # experiments: extra -D flags
        cmd += os.environ["ALAD_NVCC_EXTRA"].split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
