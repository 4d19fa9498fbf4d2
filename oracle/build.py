"""Build recipe for the oracle (test infrastructure): C restatements -> oracle/_build/, and - only
where /root/reference exists - the reference's own kernel string -> oracle/_ref/ (see ref_harness.py)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build_oracle(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("refine_oracle.c", "assoc_oracle.c") if os.path.isfile(os.path.join(HERE, f))]
    if not force and os.path.isfile(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-fopenmp", "-o", out] + srcs + ["-lm"]
    subprocess.run(cmd, check=True)
    return out


def build_ref(force: bool = False):
    from . import ref_harness as rh
    if not rh.reference_available():
        return rh.REF_KERNEL_SO if os.path.isfile(rh.REF_KERNEL_SO) else None
    if force or not os.path.isfile(rh.REF_KERNEL_SO):
        rh.build_ref_kernel()
    return rh.REF_KERNEL_SO


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_ref(force=True))
