"""In-tree build of the C-ABI CUDA core (liblidar_b200.so) for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU container; the built .so is git-ignored
but travels to the GPU box with the repo snapshot.  Only sm_100a is targeted — there is no
multi-architecture fallback by design.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "liblidar_b200.so"
OBJ_DIR = PKG_DIR / "csrc" / "_obj"
TORCH_EXT_SRC = PKG_DIR / "csrc_torch" / "lidar_torch_ext.cpp"
TORCH_EXT_PATH = PKG_DIR / "lidar_b200_torch.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # no silent a*b+c contraction anywhere: integer decisions are bit-exact
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
# kernels whose bulk fp32 math wants FMA opt back in per file
FMAD_OK = {"sa_mlp.cu"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the lidar_b200 CUDA core cannot be built")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h")):
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, verbose: bool) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    obj = OBJ_DIR / (src.stem + ".o")
    stamp_file = OBJ_DIR / (src.stem + ".stamp")
    stamp = _stamp(src)
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    flags = list(NVCC_FLAGS)
    if src.name in FMAD_OK:
        flags = [f for f in flags if f != "--fmad=false"]
    cmd = [nvcc, *flags, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    stamp_file.write_text(stamp)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link liblidar_b200.so in-tree."""
    nvcc = _nvcc()
    if force and OBJ_DIR.exists():
        shutil.rmtree(OBJ_DIR)
    srcs = _sources()
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if (not LIB_PATH.exists()) or LIB_PATH.stat().st_mtime < newest or force:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
               *map(str, objs), "--cudart", "shared",
               "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    build_torch_extension(force=force)
    return LIB_PATH


def build_torch_extension(force: bool = False) -> Path:
    """The thin PyTorch C++ extension (csrc_torch/lidar_torch_ext.cpp -> lidar_b200_torch.so, in-tree): host C++ only,
    it calls the C ABI of liblidar_b200.so and takes the current CUDA stream from PyTorch.  Compiled with g++ against
    the headers / libraries of the installed torch (same C++11 ABI setting)."""
    import torch
    tdir = Path(torch.__file__).resolve().parent
    stamp_file = OBJ_DIR / "torch_ext.stamp"
    h = hashlib.sha256()
    h.update(TORCH_EXT_SRC.read_bytes())
    for hdr in sorted(INCLUDE.glob("*.h")):
        h.update(hdr.read_bytes())
    h.update(torch.__version__.encode())
    stamp = h.hexdigest()
    if (not force and TORCH_EXT_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == stamp
            and TORCH_EXT_PATH.stat().st_mtime >= LIB_PATH.stat().st_mtime - 1e9):
        return TORCH_EXT_PATH
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    cxx = os.environ.get("CXX") or shutil.which("g++")
    if not cxx:
        raise RuntimeError("g++ not found: the torch extension cannot be built")
    cuda_home = Path(_nvcc()).resolve().parent.parent
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared",
           f"-D_GLIBCXX_USE_CXX11_ABI={1 if torch.compiled_with_cxx11_abi() else 0}",
           f"-I{tdir / 'include'}", f"-I{tdir / 'include' / 'torch' / 'csrc' / 'api' / 'include'}",
           f"-I{cuda_home / 'include'}", f"-I{INCLUDE}", str(TORCH_EXT_SRC), "-o", str(TORCH_EXT_PATH),
           f"-L{tdir / 'lib'}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
           f"-L{PKG_DIR}", "-l:liblidar_b200.so", f"-L{cuda_home / 'lib64'}", "-lcudart",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tdir / 'lib'}", f"-Wl,-rpath,{cuda_home / 'lib64'}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"g++ failed for {TORCH_EXT_SRC.name}:\n{res.stdout}\n{res.stderr}")
    stamp_file.write_text(stamp)
    return TORCH_EXT_PATH


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
    print(build_torch_extension(force="-f" in sys.argv))
