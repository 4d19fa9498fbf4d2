"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* reference under import shims.

Usable only where `/root/reference` exists (the build container); never on the
GPU box and never from the product path.  It is how the oracle in this
directory is pinned: golden vectors under
`tests/golden/` are produced by running the reference's own code through this
module (`tests/golden/make_golden.py`).

Shims (nothing from the reference is copied into the repo):
  * `matplotlib`, `matplotlib.pyplot`, `matplotlib.cm`: empty stub modules (the
    import at instances.py:14 is unused).
  * `pycuda.*`: a fake whose `SourceModule` takes the reference's CUDA kernel
    *string* (box_fusion.py:63-407), prefixes host definitions of
    `__device__/__global__/blockIdx/threadIdx/atomicAdd_system`, drops the unused
    `<curand_kernel.h>` include and pipes the result through `g++` (stdin, no
    copy on disk) into `oracle/_ref/libref_kernel.so`.  `cuda.In/InOut` are numpy
    pass-throughs.  The reference's Python optimiser loop then runs unmodified.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("BOXFUSION_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(_HERE, "_ref")
REF_KERNEL_SO = os.path.join(REF_OUT, "libref_kernel.so")

_PREFIX = r"""
#include <cmath>
#include <cstdlib>
#include <math.h>
#include <stdlib.h>
#include <algorithm>
using std::max; using std::min; using std::abs;
#define __device__
#define __global__
struct bf_dim3 { unsigned x, y, z; };
static thread_local bf_dim3 blockDim, blockIdx, threadIdx;
template <class T, class U> static inline void atomicAdd_system(T* p, U v) { *p += v; }
"""

_SUFFIX = r"""
extern "C" void bf_ref_run_grid(int gx, int gy, int bx,
        float* a0, float* a1, float* a2, float* a3, float* a4, float* a5,
        float* a6, float* a7, float* a8, float* a9, float* a10) {
    blockDim.x = bx; blockDim.y = 1; blockDim.z = 1;
    for (int by = 0; by < gy; ++by)           // views ascending: the host order of the view sum
        for (int b = 0; b < gx; ++b)
            for (int t = 0; t < bx; ++t) {
                blockIdx.x = b; blockIdx.y = by; blockIdx.z = 0;
                threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
                compute_iou_value(a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10);
            }
}
"""


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "boxfusion", "box_fusion.py"))


def extract_kernel_string() -> str:
    """The CUDA source held as a Python string literal in box_fusion.py:63-407."""
    path = os.path.join(REFERENCE_ROOT, "boxfusion", "box_fusion.py")
    text = open(path, "r").read()
    start = text.index('SourceModule("""') + len('SourceModule("""')
    end = text.index('"""', start)
    return text[start:end]


def compile_kernel_string(src: str, out_so: str = REF_KERNEL_SO) -> str:
    """g++ the kernel string as host C++ (stdin -> .so); outputs only under oracle/_ref/."""
    os.makedirs(os.path.dirname(out_so), exist_ok=True)
    body = src.replace("#include <curand_kernel.h>", "")
    full = _PREFIX + body + _SUFFIX
    cmd = ["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", "-", "-o", out_so]
    r = subprocess.run(cmd, input=full.encode(), capture_output=True)
    if r.returncode != 0:
        raise RuntimeError("host compile of the reference kernel string failed:\n" + r.stderr.decode())
    return out_so


class RefKernel:
    """ctypes view of oracle/_ref/libref_kernel.so (the reference kernel, host-compiled)."""

    def __init__(self, so_path: str = REF_KERNEL_SO):
        self.lib = ctypes.CDLL(so_path)
        fp = ctypes.POINTER(ctypes.c_float)
        self.lib.bf_ref_run_grid.argtypes = [ctypes.c_int] * 3 + [fp] * 11
        self.lib.bf_ref_run_grid.restype = None

    def launch(self, arrays, block, grid):
        fp = ctypes.POINTER(ctypes.c_float)
        ptrs = [a.ctypes.data_as(fp) for a in arrays]
        self.lib.bf_ref_run_grid(int(grid[0]), int(grid[1]), int(block[0]), *ptrs)


# ---------------------------------------------------------------------------
# fake pycuda
# ---------------------------------------------------------------------------

class _Arg:
    def __init__(self, arr, writeback):
        self.src = arr
        self.arr = np.ascontiguousarray(arr, dtype=np.float32)
        self.writeback = writeback


class _Function:
    def __init__(self, kern: RefKernel):
        self.kern = kern

    def __call__(self, *args, block, grid):
        self.kern.launch([a.arr for a in args], block, grid)
        for a in args:
            if a.writeback and a.arr is not a.src:
                a.src[...] = a.arr


class _SourceModule:
    def __init__(self, src, no_extern_c=False, **kw):
        compile_kernel_string(src)
        self.kern = RefKernel()

    def get_function(self, name):
        assert name == "compute_iou_value"
        return _Function(self.kern)


def _install_shims():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    pycuda = types.ModuleType("pycuda")
    drv = types.ModuleType("pycuda.driver")

    class PointerHolderBase:  # box_fusion.py:19 subclasses it at import time
        pass

    drv.PointerHolderBase = PointerHolderBase
    drv.In = lambda a: _Arg(a, False)
    drv.InOut = lambda a: _Arg(a, True)
    comp = types.ModuleType("pycuda.compiler")
    comp.SourceModule = _SourceModule
    auto = types.ModuleType("pycuda.autoprimaryctx")
    gpa = types.ModuleType("pycuda.gpuarray")
    pycuda.driver, pycuda.compiler, pycuda.autoprimaryctx, pycuda.gpuarray = drv, comp, auto, gpa
    sys.modules.update({"pycuda": pycuda, "pycuda.driver": drv, "pycuda.compiler": comp,
                        "pycuda.autoprimaryctx": auto, "pycuda.gpuarray": gpa})


_REF = None


def load_reference():
    """Import the reference's hot-path modules; returns a namespace of its own classes."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_shims()
    # make sure `boxfusion` resolves to the reference, not to our drop-in alias
    for k in [k for k in sys.modules if k == "boxfusion" or k.startswith("boxfusion.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import boxfusion.boxes as boxes
        import boxfusion.instances as instances
        import boxfusion.box_manager as box_manager
        import boxfusion.box_fusion as box_fusion
    finally:
        sys.path.remove(REFERENCE_ROOT)
    ns = types.SimpleNamespace(
        boxes=boxes, instances=instances, box_manager=box_manager, box_fusion=box_fusion,
        Instances3D=instances.Instances3D, GeneralInstance3DBoxes=boxes.GeneralInstance3DBoxes,
        BoxManager=box_manager.BoxManager, BoxFusion=box_fusion.BoxFusion,
        nms_3d=instances.nms_3d, calculate_obb_iou=instances.calculate_obb_iou)
    _REF = ns
    return ns


class _Stub(types.ModuleType):
    """Import-time stand-in for the viewer stack tools/utils.py pulls in (rerun, open3d); never called by what we use."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return None


_TOOLS = None


def load_reference_tools():
    """The reference's tools/utils.py (post_process :302-317, save_box :322-332, load_data :335-340), unmodified, with the
    viewer modules it imports at the top (rerun, open3d - absent here) stubbed."""
    global _TOOLS
    if _TOOLS is not None:
        return _TOOLS
    load_reference()
    for name in ("rerun", "rerun.blueprint", "open3d", "torchvision", "torchvision.transforms", "torchvision.transforms.functional"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    for k in [k for k in sys.modules if k == "tools" or k.startswith("tools.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import tools.utils as tu
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _TOOLS = tu
    return tu


def build_ref_kernel() -> str:
    """Recipe entry (called by oracle/build.py): compile the kernel string into oracle/_ref/."""
    return compile_kernel_string(extract_kernel_string())


if __name__ == "__main__":
    print(build_ref_kernel())
