#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Build the reference's own CUDA kernels as the GPU oracle.

Compiles /root/reference/models/csrc/{intersection,raymarching,volumerendering,losses}.cu
and binding.cpp *in place* (no source is copied or edited) for sm_100 into the
python extension module ``oracle/_ref/vren_ref.so`` (git-ignored, but shipped to
the GPU box by gpurun).  The only addition is the force-included
``oracle/ref_compat/vren_ref_compat.h`` (see its header for why).  Flags mirror
what the reference's setup.py yields through torch's BuildExtension
(models/csrc/setup.py:22-28: ``-O2`` only => nvcc default -fmad=true, IEEE div/sqrt),
because marching bit-exactness depends on the same FMA contraction.

The reference's own build system (setup.py) is NOT run.  On a box without
/root/reference (the GPU box) this script is a no-op: the prebuilt .so is used.

The same route ships the reference's UNCHANGED Python hot-path files (models/{__init__,custom_functions,rendering,ngp_mt}.py,
losses.py, datasets/hypersim_src/utils.py) into git-ignored ``oracle/_ref/py/`` (`stage_py`), so that the GPU box can run
the reference's own render() / NGPMT / NeRFMTLoss on top of the drop-in shims (tests/test_reference_on_shims_gpu.py,
bench.py's cpu_baseline leg).  The files are byte-for-byte copies made at build time, never committed; only the two package
markers ``datasets/__init__.py`` / ``datasets/hypersim_src/__init__.py`` are written empty (the reference's datasets/__init__.py
imports every dataset class and with them cv2 / h5py / kornia, none of which the hot path uses).

Usage:  python oracle/build_ref.py [--force] [--ptx]
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NCN_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "models", "csrc")
OUT = os.path.join(HERE, "_ref")
BUILD = os.path.join(HERE, "_build")
COMPAT = os.path.join(HERE, "ref_compat", "vren_ref_compat.h")
NAME = "vren_ref"
CU = ["intersection.cu", "raymarching.cu", "volumerendering.cu", "losses.cu"]
CPP = ["binding.cpp"]


def _torch_paths():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths()
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    return inc, lib


def available():
    return os.path.isdir(SRC)


def target():
    return os.path.join(OUT, NAME + ".so")


def build(force=False, ptx=False, verbose=True):
    if not available():
        if verbose:
            print(f"[oracle/build_ref] {SRC} not present: using prebuilt {target()} if any")
        return os.path.exists(target())
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    stage_py(verbose=False)
    srcs = [os.path.join(SRC, f) for f in CU + CPP]
    newest = max(os.path.getmtime(p) for p in srcs + [COMPAT, __file__])
    if not force and os.path.exists(target()) and os.path.getmtime(target()) >= newest:
        if verbose:
            print(f"[oracle/build_ref] up to date: {target()}")
        return True
    inc, lib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    common = [f"-I{os.path.join(SRC, 'include')}"] + [f"-I{i}" for i in inc] + [f"-I{pyinc}"]
    defs = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=1"]
    nv_defs = ["-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
               "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"]
    jobs = []
    for f in CU:
        o = os.path.join(BUILD, f + ".o")
        jobs.append((o, ["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr",
                         "-gencode", "arch=compute_100,code=sm_100",
                         "-Xcompiler", "-fPIC", "-w", "-include", COMPAT]
                     + common + defs + nv_defs + ["-c", os.path.join(SRC, f), "-o", o]))
        if ptx:
            p = os.path.join(BUILD, f + ".ptx")
            jobs.append((p, ["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr",
                             "-gencode", "arch=compute_100,code=compute_100", "-w",
                             "-include", COMPAT] + common + defs + nv_defs
                         + ["-ptx", os.path.join(SRC, f), "-o", p]))
    for f in CPP:
        o = os.path.join(BUILD, f + ".o")
        jobs.append((o, ["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-include", COMPAT,
                         "-I/usr/local/cuda/include"] + common + defs
                     + ["-c", os.path.join(SRC, f), "-o", o]))

    def run(job):
        out, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"[oracle/build_ref] failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return out

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(BUILD, f + ".o") for f in CU + CPP]
    link = ["g++", "-shared", "-o", target()] + objs + [
        f"-L{lib}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
        "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{lib}"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"[oracle/build_ref] link failed\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[oracle/build_ref] built {target()}")
    return True


PY_OUT = os.path.join(OUT, "py")
PY_FILES = ["models/__init__.py", "models/custom_functions.py", "models/rendering.py", "models/ngp_mt.py", "losses.py",
            "datasets/hypersim_src/utils.py"]
PY_MARKERS = ["datasets/__init__.py", "datasets/hypersim_src/__init__.py"]


def stage_py(verbose=True):
    """byte-for-byte copies of the reference's hot-path Python files -> oracle/_ref/py/ (git-ignored, travels with gpurun)"""
    import shutil
    if not os.path.isdir(os.path.join(REF, "models")):
        ok = all(os.path.exists(os.path.join(PY_OUT, f)) for f in PY_FILES)
        if verbose:
            print(f"[oracle/build_ref] {REF} not present: staged python files {'found' if ok else 'MISSING'} in {PY_OUT}")
        return ok
    for f in PY_FILES:
        dst = os.path.join(PY_OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        src = os.path.join(REF, f)
        if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
            shutil.copyfile(src, dst)
    for f in PY_MARKERS:
        dst = os.path.join(PY_OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").close()
    return True


def staged_py():
    """path of the staged reference python tree, or None"""
    return PY_OUT if all(os.path.exists(os.path.join(PY_OUT, f)) for f in PY_FILES) else None


def load():
    """Import the prebuilt reference extension (GPU box: tests only)."""
    import importlib.util
    import torch  # noqa: F401  (loads libtorch first)
    if not os.path.exists(target()):
        return None
    spec = importlib.util.spec_from_file_location(NAME, target())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, ptx="--ptx" in sys.argv)
    sys.exit(0 if ok else 1)
