"""Kernel-tuning builds: compile bf_refine.cu with extra -D flags into boxfusion_b200/lib/variants/lib_<name>.so
(the other objects are shared with the production build).  Select one at run time with BOXFUSION_B200_LIB=<path>.

    python tools/build_variants.py base="" slots256="-DBF_CNT_SLOTS=256" ...   # name=extra nvcc flags (any -D the sources honour)

The round-1 sweep (profiles/r1_v10_variant_shape_sweep.txt) used this with the then compile-time switches for rolled loops and
register caps; the winners became the three template instantiations of bf_refine_kernel.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from boxfusion_b200 import build as B   # noqa: E402


def main():
    B.build()
    out_dir = os.path.join(B.OUT_DIR, "variants")
    os.makedirs(out_dir, exist_ok=True)
    others = [os.path.join(B.OUT_DIR, s.replace(".cu", ".o")) for s in B.SOURCES if s != "bf_refine.cu"]
    for arg in sys.argv[1:]:
        name, flags = arg.split("=", 1)
        obj = os.path.join(out_dir, f"bf_refine_{name}.o")
        cmd = ["nvcc"] + B.ARCH + B.COMMON + B.SOURCES["bf_refine.cu"] + flags.split() + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, "bf_refine.cu"), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            raise SystemExit(1)
        lines = (r.stdout + r.stderr).splitlines()
        for i, ln in enumerate(lines):
            if "Compiling entry function '_Z16bf_refine_kernel" in ln:
                print(name, "|", lines[i + 2].strip(), "|", lines[i + 3].strip() if i + 3 < len(lines) else "")
        lib = os.path.join(out_dir, f"lib_{name}.so")
        subprocess.run(["nvcc"] + B.ARCH + ["-shared", "-o", lib, obj] + others + ["-lcudart"], check=True)


if __name__ == "__main__":
    main()
