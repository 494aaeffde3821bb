"""Build recipes for the checkers -- TEST INFRASTRUCTURE ONLY.

build_c_oracle(): gcc oracle/mas_oracle.c -> oracle/_build/libmas_oracle.so
build_ref():      cythonize the REFERENCE's own /root/reference/model/monotonic_align/core.pyx,
                  from where it lies, into oracle/_ref/ (git-ignored; outputs only, no reference
                  sources are copied into the repo) and compile it with gcc.  Only possible in the
                  build container; on the GPU box the prebuilt oracle/_ref/*.so travels along.
"""
import os
import subprocess
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/model/monotonic_align/core.pyx"


def build_c_oracle(force=False):
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(_HERE, "mas_oracle.c")
    out = os.path.join(out_dir, "libmas_oracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src])
    return out


def ref_so_path():
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.join(_HERE, "_ref", "core" + ext)


def build_ref(force=False):
    """Returns the path of the compiled reference MAS module, or None if /root/reference is absent."""
    out = ref_so_path()
    if os.path.exists(out) and not force:
        return out
    if not os.path.exists(REF_PYX):
        return None
    import numpy
    out_dir = os.path.join(_HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    c_file = os.path.join(out_dir, "core.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fwrapv", "-fno-strict-aliasing",
                           "-I", inc, "-I", numpy.get_include(), "-o", out, c_file])
    return out


def load_ref():
    """Import oracle/_ref/core*.so (the compiled reference) or return None."""
    import importlib.util
    p = ref_so_path()
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location("core", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_c_oracle(force=True))
    print(build_ref(force=True))
