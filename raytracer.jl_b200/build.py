"""Compile librt_sssp.so (all CUDA sources under csrc/) for sm_100a with nvcc.  In-tree output so that the
shared object travels with the repository snapshot to the GPU box."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "librt_sssp.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "-fmad=false",          # the reference never contracts a*b+c (plain Julia): keep fp64 bit-exact
    "-I" + os.path.join(ROOT, "include"),
]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = (sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + glob.glob(os.path.join(HERE, "csrc", "*.h"))
            + [os.path.join(ROOT, "include", "rt_sssp.h")])
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Kernel-variant A/B builds: build_variants/<name>/librt_sssp.so with extra -D flags (load it with RT_SSSP_LIB)."""
    vdir = os.path.join(ROOT, "build_variants", name)
    os.makedirs(vdir, exist_ok=True)
    so = os.path.join(vdir, "librt_sssp.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(vdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, "-ccbin", ccbin] + NVCC_FLAGS + list(defines) + ["-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on " + src)
    subprocess.check_call([nvcc, "-ccbin", ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs +
                          ["-lcudart", "-ldl"])
    return so


def build(force=False, verbose=False):
    if os.environ.get("RT_SSSP_LIB"):  # a prebuilt kernel variant is selected: leave the default library alone
        return os.environ["RT_SSSP_LIB"]
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, "-ccbin", ccbin] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out.decode())
        if p.returncode:
            raise RuntimeError("nvcc failed on " + src)
    cmd = [nvcc, "-ccbin", ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
