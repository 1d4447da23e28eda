"""nvcc builds of the core library and of per-weak-form physics modules (sm_100a, in-tree).

`libechem_b200.so` is built once; a physics module is built per distinct
generated header (content-hashed, cached under echemfem_b200/lib/modules/).
Both are git-ignored build artefacts that travel with the tree.
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIBDIR = os.path.join(HERE, "lib")
MODDIR = os.path.join(LIBDIR, "modules")
CORE_LIB = os.path.join(LIBDIR, "libechem_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-shared", "-Xcompiler", "-fPIC", "-I" + INCLUDE, "-I" + CSRC]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), res.stdout, res.stderr))
    return res


def build_core(force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, "ech_api.cu"), os.path.join(INCLUDE, "echem_b200.h"),
            os.path.join(INCLUDE, "echem_b200_module.h")]
    if not force and _newer(CORE_LIB, srcs):
        return CORE_LIB
    _run([_nvcc()] + NVCC_FLAGS + [srcs[0], "-o", CORE_LIB, "-ldl"])
    return CORE_LIB


def skeleton_hash():
    h = hashlib.sha1()
    for name in ("ech_dg.cuh", "ech_dg_coop.cuh", "ech_dg_coop2.cuh", "ech_cg.cuh", "ech_module.cu"):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    with open(os.path.join(INCLUDE, "echem_b200_module.h"), "rb") as f:
        h.update(f.read())
    return h.hexdigest()[:8]


def _prune_stale():
    """Drop modules built against an older kernel skeleton (they would never be loaded again)."""
    tag = "_" + skeleton_hash()
    for name in os.listdir(MODDIR):
        if name.startswith("libech_phys_") and name.endswith(".so") and tag not in name:
            try:
                os.remove(os.path.join(MODDIR, name))
            except OSError:
                pass


def module_path(key, extra_flags=()):
    tag = skeleton_hash()
    if extra_flags:
        tag += "_" + hashlib.sha1(" ".join(extra_flags).encode()).hexdigest()[:6]
    return os.path.join(MODDIR, "libech_phys_%s_%s.so" % (key, tag))


def build_module(header_text, key, force=False, extra_flags=()):
    """Compile one physics module; returns the path of the shared object."""
    os.makedirs(MODDIR, exist_ok=True)
    _prune_stale()
    extra_flags = tuple(extra_flags) or tuple(os.environ.get("ECH_MODULE_FLAGS", "").split())
    out = module_path(key, extra_flags)
    if os.path.exists(out) and not force:
        return out
    hdr = os.path.join(MODDIR, "ech_gen_%s.cuh" % key)
    with open(hdr, "w") as f:
        f.write(header_text)
    tmp = out + ".tmp.%d" % os.getpid()
    # physics modules are dlopen'ed side by side and define the same C++ names: no STB_GNU_UNIQUE
    # symbols (inline-function statics would otherwise be shared between modules)
    _run([_nvcc()] + NVCC_FLAGS + ["-Xcompiler", "-fno-gnu-unique", "-Xcompiler", "-fvisibility=hidden"] +
         list(extra_flags) + ["-include", hdr, os.path.join(CSRC, "ech_module.cu"), "-o", tmp])
    os.replace(tmp, out)
    return out
