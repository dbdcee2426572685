"""Builds libmprg.so (sm_100a only) in-tree with nvcc; used by __graft_entry__.build()."""
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmprg.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]
# the KMeans kernel restates scikit-learn's float64 operation order: no FMA contraction there
PER_FILE = {"kmeans.cu": ["-fmad=false"]}
# kmeans.cu is compiled a second time for deep loci: every initialisation on a group of CTAs that meet
# at a global-memory barrier; -dlcm=cg makes every global load of that object an L2 (coherent) load
EXTRA_OBJECTS = [("kmeans.cu", "kmeans_group", ["-fmad=false", "-DMPRG_KM_GROUP", "-Xptxas", "-dlcm=cg"])]


def sources():
    # *.cu: kernels + engine; *.cpp: host-only code (loader / writers), compiled by nvcc's host compiler
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def compile_units():
    units = [(src, src.stem, PER_FILE.get(src.name, [])) for src in sources()]
    units += [(CSRC / name, stem, flags) for name, stem, flags in EXTRA_OBJECTS]
    return units


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "mprg.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force=False, verbose=False, extra_flags=(), lib_out=None):
    """extra_flags / lib_out: experiment builds (e.g. -DMPRG_SCAN_MIN_BLOCKS=8 into libmprg_x.so, picked up
    through the MPRG_LIB environment variable); the default build is what ships."""
    variant = bool(extra_flags) or lib_out is not None
    if not variant and not force and not needs_build():
        return LIB
    objdir = CSRC / ("build_variant" if variant else "build")
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src, stem, flags in compile_units():
        obj = objdir / (stem + ".o")
        cmd = [NVCC, *ARCH, *COMMON, *flags, *extra_flags, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((stem, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libmprg build failed")
    lib = Path(lib_out) if lib_out is not None else LIB
    cmd = [NVCC, *ARCH, "-shared", "-o", str(lib), *objs, "-lz"]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[len("--out="):] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=extra,
                        lib_out=outs[0] if outs else None))
