"""In-tree build of the two native artefacts (no JIT cache, so the .so files travel with gpurun):

  directvoxgo_b200/libdvgo_b200.so   nvcc, sm_100a only, CUDA headers only  -> the C ABI
  directvoxgo_b200/_C<ext>.so        g++,  torch headers                    -> the thin torch binding

`python directvoxgo_b200/build.py [--force]` (run as a script: the package itself refuses to import
before these exist); also called by __graft_entry__.build().
"""
import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]

LIB_NAME = "libdvgo_b200.so"
EXT_NAME = "_C" + sysconfig.get_config_var("EXT_SUFFIX")


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return p.stdout


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(HERE, "..", "include")
    hs += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h")]
    return hs


def build_lib(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    cus = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = _headers()
    out = os.path.join(HERE, LIB_NAME)

    def compile_one(cu):
        obj = os.path.join(BUILD, os.path.basename(cu) + ".o")
        stamp = obj + ".sha"
        dig = _digest([cu] + hdrs, " ".join(NVCC_FLAGS))
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            return obj, False
        log = _run([NVCC] + NVCC_FLAGS + ["-c", cu, "-o", obj], obj + ".log")
        if verbose:
            print(log)
        open(stamp, "w").write(dig)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(8, len(cus))) as ex:
        res = list(ex.map(compile_one, cus))
    objs = [r[0] for r in res]
    if force or any(r[1] for r in res) or not os.path.exists(out):
        _run([NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
              "-lcudart", "-Xlinker", "-rpath," + os.path.join(CUDA_HOME, "lib64")],
             os.path.join(BUILD, "link_lib.log"))
    return out


def build_binding(force=False):
    import torch
    os.makedirs(BUILD, exist_ok=True)
    tdir = os.path.dirname(torch.__file__)
    cpps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cpp"))
    hdrs = _headers()
    out = os.path.join(HERE, EXT_NAME)
    flags = [
        "-O2", "-fPIC", "-std=c++17", "-w", "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-DTORCH_EXTENSION_NAME=_C", "-D_GLIBCXX_USE_CXX11_ABI=1",
        "-I" + os.path.join(tdir, "include"),
        "-I" + os.path.join(tdir, "include", "torch", "csrc", "api", "include"),
        "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(CUDA_HOME, "include"),
    ]

    def compile_one(cpp):
        obj = os.path.join(BUILD, os.path.basename(cpp) + ".o")
        stamp = obj + ".sha"
        dig = _digest([cpp] + hdrs, " ".join(flags) + torch.__version__)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            return obj, False
        _run(["g++"] + flags + ["-c", cpp, "-o", obj], obj + ".log")
        open(stamp, "w").write(dig)
        return obj, True

    with ThreadPoolExecutor(max_workers=4) as ex:
        res = list(ex.map(compile_one, cpps))
    objs = [r[0] for r in res]
    lib = os.path.join(HERE, LIB_NAME)
    if force or any(r[1] for r in res) or not os.path.exists(out) or \
            os.path.getmtime(lib) > os.path.getmtime(out):
        _run(["g++", "-shared", "-o", out] + objs + [
            "-L" + HERE, "-ldvgo_b200", "-Wl,-rpath,$ORIGIN",
            "-L" + os.path.join(tdir, "lib"), "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
            "-ltorch", "-ltorch_python", "-Wl,-rpath," + os.path.join(tdir, "lib"),
            "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart",
        ], os.path.join(BUILD, "link_ext.log"))
    return out


def build_all(force=False, verbose=False):
    lib = build_lib(force=force, verbose=verbose)
    ext = build_binding(force=force)
    return lib, ext


if __name__ == "__main__":
    force = "--force" in sys.argv
    lib, ext = build_all(force=force, verbose="-v" in sys.argv)
    print("built", lib)
    print("built", ext)
