"""Build libptf_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libptf_b200.so")
SOURCES = ["ptf_api.cu", "engine_cufft.cu", "engine_fused.cu", "engine_fused3d.cu", "engine_fused1d.cu", "engine_mqg.cu",
           "engine_slab2d.cu", "expr_flow.cu"]
FUSED_SIZES = [64, 128, 256, 512, 1024, 2048, 4096]   # fused_inst.cu is compiled once per transform length (in parallel)
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ptf_b200.h"),
                                                                 os.path.abspath(__file__)]
    newest = _newest(deps)
    if newest > os.path.getmtime(LIB):
        return True
    # every object must also be newer than its own source and every header (a file edited WHILE a build was running
    # leaves a fresh-looking library linked from a stale object)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    jobs = [(s, s.replace(".cu", ".o")) for s in SOURCES] + [("fused_inst.cu", f"fused_inst_{n}.o") for n in FUSED_SIZES]
    for src, o in jobs:
        po = os.path.join(HERE, "build", o)
        if not os.path.exists(po) or os.path.getmtime(po) < _newest(hdrs + [os.path.join(CSRC, src)]):
            return True
    return False


def build(force=False, verbose=False, with_nccl=True, variant=None, defines=()):
    """variant / defines: developer knob for A/B experiments — builds libptf_b200_<variant>.so with extra -D flags into
    its own object directory; load it with PTF_LIB_PATH=<that file> (see _capi.load)."""
    global LIB
    if variant:
        return _build(True, verbose, with_nccl, os.path.join(HERE, f"libptf_b200_{variant}.so"),
                      os.path.join(HERE, f"build_{variant}"), [f"-D{d}" for d in defines])
    if not force and not needs_build():
        return LIB
    return _build(force, verbose, with_nccl, LIB, os.path.join(HERE, "build"), [])


def _build(force, verbose, with_nccl, LIB, BUILD, extra_defs):
    nvcc = _nvcc()
    objs = []
    common = [nvcc, "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
              "-I", os.path.join(HERE, "..", "include"), *extra_defs]
    if verbose:
        common += ["-Xptxas", "-v"]
    if with_nccl:
        common += ["-DPTF_WITH_NCCL"]
    os.makedirs(BUILD, exist_ok=True)
    procs = []
    jobs = [(s, s.replace(".cu", ".o"), []) for s in SOURCES]
    jobs += [("fused_inst.cu", f"fused_inst_{n}.o", [f"-DPTF_INST_N={n}"]) for n in FUSED_SIZES]
    for s, oname, extra in jobs:
        o = os.path.join(BUILD, oname)
        objs.append(o)
        src = os.path.join(CSRC, s)
        if (not force) and os.path.exists(o) and os.path.getmtime(o) > _newest(
                [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [src]):
            continue
        procs.append((oname, subprocess.Popen(common + extra + ["-c", src, "-o", o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    link = [nvcc, "-shared", *ARCH, "-o", LIB, *objs, "-L/usr/local/cuda/lib64", "-lcufft", "-lnvrtc",
            "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    if with_nccl:
        link += ["-lnccl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:   # python build.py --variant nb8 PTF_RK4_NB=8
        i = sys.argv.index("--variant")
        print(build(variant=sys.argv[i + 1], defines=[a for a in sys.argv[i + 2:] if "=" in a or a.isidentifier()]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
