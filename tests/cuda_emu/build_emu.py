"""TEST INFRASTRUCTURE ONLY: compile a CUDA-core translation unit of aladin_b200/csrc for the host-thread emulator
(tests/cuda_emu/cuda_emu.h).  The source is used as it is, except for two textual rewrites g++ cannot parse:
`kernel<<<grid, block, smem, stream>>>(args);` and `extern __shared__ T name[];`."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "aladin_b200", "csrc")
OUT = os.path.join(HERE, "_build")

_LAUNCH = re.compile(r"(\b[\w:]+(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)
_DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")


def translate(src):
    src, n = _LAUNCH.subn(lambda m: f"cuda_emu::launch_cfg({m.group(2)}, [&] {{ {m.group(1)}({m.group(3)}); }});", src)
    src = _DYN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = static_cast<{m.group(1)}*>(cuda_emu::dynamic_smem());", src)
    return src, n


def tsan_runtime():
    """Path of gcc's libtsan.so (to LD_PRELOAD into the python that loads a tsan=True build), or None."""
    res = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True)
    path = res.stdout.strip()
    return path if res.returncode == 0 and os.path.isabs(path) and os.path.exists(path) else None


def build(name, tsan=False):
    """aladin_b200/csrc/<name>.cu -> tests/cuda_emu/_build/lib<name>_emu.so (rebuilt when the source changed).
    tsan=True builds lib<name>_emu_tsan.so with -fsanitize=thread: every CUDA thread is a host thread and the
    barriers are pthread barriers, so a missing __syncthreads / __syncwarp between a shared-memory write and a
    read by another thread is reported as a data race."""
    os.makedirs(OUT, exist_ok=True)
    cu = os.path.join(CSRC, name + ".cu")
    cpp = os.path.join(OUT, name + "_emu.cpp")
    lib = os.path.join(OUT, f"lib{name}_emu{'_tsan' if tsan else ''}.so")
    deps = [cu, os.path.join(HERE, "cuda_emu.h"), os.path.join(CSRC, "common.h"), os.path.join(ROOT, "include", "alad_b200.h"),
            os.path.abspath(__file__)]
    if os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    text, n = translate(open(cu).read())
    with open(cpp, "w") as f:
        f.write('#include "cuda_emu.h"\n' + text)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-DALAD_CPU_EMU", "-shared", "-fPIC", "-pthread", "-w", "-Wl,-Bsymbolic", "-I", HERE, "-I", CSRC,
           "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, "-o", lib, cpp]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stderr[-4000:])
    return lib


FULL_LIBRARY = ["cabi", "pack", "fused", "mrsw_bwd", "losses", "train_step", "distill", "misc_sim", "scan_pool", "rank", "retrieval", "pairs", "peer", "h2d"]


def build_library(tsan=False):
    """Every translation unit of the C-ABI library except the tcgen05 kernel (csrc/mrsw_fwd.cu), whose entry point
    alad_mrsw_scores_fwd is provided by the CPU double tests/cuda_emu/mrsw_fwd_double.cpp -> one emulated
    libalad_b200_emu.so with the complete symbol set of include/alad_b200.h.  With it the native compositions
    (alad_scores_fused, alad_train_losses_fwd/_bwd: host bookkeeping, tile tables, workspace layout, launch order)
    and the Python host layer above them run on CPU tensors."""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, f"libalad_b200_emu{'_tsan' if tsan else ''}.so")
    double = os.path.join(HERE, "mrsw_fwd_double.cpp")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"), double,
                                                                 os.path.join(ROOT, "include", "alad_b200.h"), os.path.abspath(__file__)]
    if os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    flags = ["-std=c++17", "-O1", "-g", "-DALAD_CPU_EMU", "-DCUDA_EMU_FULL_LIBRARY", "-fPIC", "-pthread", "-w", "-I", HERE, "-I", CSRC,
             "-I", os.path.join(ROOT, "include"), "-I", cuda_inc] + (["-fsanitize=thread"] if tsan else [])
    objs, procs = [], []
    for name in FULL_LIBRARY:
        text, _ = translate(open(os.path.join(CSRC, name + ".cu")).read())
        cpp = os.path.join(OUT, f"{name}_lib{'_tsan' if tsan else ''}.cpp")
        with open(cpp, "w") as f:
            f.write('#include "cuda_emu.h"\n' + text)
        obj = cpp[:-4] + ".o"
        objs.append(obj)
        procs.append((cpp, subprocess.Popen(["g++"] + flags + ["-c", cpp, "-o", obj], stderr=subprocess.PIPE, text=True)))
    obj = os.path.join(OUT, f"mrsw_fwd_double{'_tsan' if tsan else ''}.o")
    objs.append(obj)
    procs.append((double, subprocess.Popen(["g++"] + flags + ["-c", double, "-o", obj], stderr=subprocess.PIPE, text=True)))
    for src, pr in procs:
        err = pr.communicate()[1]
        if pr.returncode != 0:
            raise RuntimeError(f"g++ failed on {src}:\n" + err[-4000:])
    res = subprocess.run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic", "-Wl,--no-undefined"] + (["-fsanitize=thread"] if tsan else []) +
                         ["-o", lib] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stderr[-4000:])
    return lib


if __name__ == "__main__":
    import sys
    print(build_library() if sys.argv[1:] == ["all"] else build(sys.argv[1] if len(sys.argv) > 1 else "scan_pool"))
