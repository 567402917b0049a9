"""Build the reference's OWN native modules into ``oracle/_ref/`` (TEST INFRASTRUCTURE).

This compiles ``deepgrp/sequence.pyx + maxcalc.c`` and ``deepgrp/_mss/pymss.pyx + _mss/mss.c``
straight from ``/root/reference`` (mirroring the two ``Extension`` entries of the reference's
``build.py:6-13``), using a scratch copy under a temp dir because the reference tree is
read-only.  Only the resulting ``.so`` files (plus two empty stub modules written by this
script) land in ``oracle/_ref/deepgrp_ref/``; no reference source is copied into the repo.

One glue edit is applied to the scratch copy: ``pymss.pyx`` line 10 ``cimport mss`` becomes
``from mss cimport msseg_t, mss_find_all`` (Cython 3 no longer star-imports a same-named .pxd).

The modules are built under the package name ``deepgrp`` (their init symbol is
``PyInit_sequence`` / ``PyInit_mss``, which does not depend on the package), and are placed in
``oracle/_ref/deepgrp/`` together with an empty ``preprocessing`` stub, because
``sequence.pyx:9`` imports ``deepgrp.preprocessing`` (which needs pandas and is unused).

Run:  python oracle/build_ref.py            (needs /root/reference; a no-op otherwise)
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DEEPGRP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF, "deepgrp", "sequence.pyx"))


def built() -> bool:
    pkg = os.path.join(OUT, "deepgrp")
    if not os.path.isdir(pkg):
        return False
    names = os.listdir(pkg)
    return any(n.startswith("sequence") and n.endswith(".so") for n in names) and \
        any(n.startswith("mss") and n.endswith(".so") for n in names)


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the two compiled reference modules."""
    if built() and not force:
        return True
    if not have_reference():
        return False
    import numpy
    scratch = tempfile.mkdtemp(prefix="deepgrp_ref_build_")
    try:
        src = os.path.join(scratch, "deepgrp")
        shutil.copytree(os.path.join(REF, "deepgrp"), src)
        pyx = os.path.join(src, "_mss", "pymss.pyx")
        with open(pyx) as fh:
            text = fh.read()
        text = text.replace("\ncimport mss\n", "\nfrom mss cimport msseg_t, mss_find_all\n")
        with open(pyx, "w") as fh:
            fh.write(text)
        setup = os.path.join(scratch, "setup_ref.py")
        with open(setup, "w") as fh:
            fh.write(
                "import numpy\n"
                "from setuptools import setup, Extension\n"
                "from Cython.Build import cythonize\n"
                "ext = [\n"
                " Extension('deepgrp.mss', ['deepgrp/_mss/pymss.pyx', 'deepgrp/_mss/mss.c'],\n"
                "   include_dirs=[numpy.get_include(), 'deepgrp', 'deepgrp/_mss']),\n"
                " Extension('deepgrp.sequence', ['deepgrp/sequence.pyx', 'deepgrp/maxcalc.c'],\n"
                "   include_dirs=[numpy.get_include(), 'deepgrp']),\n"
                "]\n"
                "setup(name='deepgrp_ref', packages=[],\n"
                "      ext_modules=cythonize(ext, language_level=3, include_path=['deepgrp']))\n")
        env = dict(os.environ)
        env.setdefault("CFLAGS", "-O2")
        subprocess.run([sys.executable, setup, "build_ext", "--inplace", "-q"],
                       cwd=scratch, check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        pkg = os.path.join(OUT, "deepgrp")
        os.makedirs(pkg, exist_ok=True)
        suffix = sysconfig.get_config_var("EXT_SUFFIX")
        for mod in ("mss", "sequence"):
            shutil.copy(os.path.join(src, mod + suffix), os.path.join(pkg, mod + suffix))
        # stubs written by us (not reference code): make `import deepgrp.preprocessing` succeed
        with open(os.path.join(pkg, "__init__.py"), "w") as fh:
            fh.write('"""oracle/_ref: compiled reference natives (built by oracle/build_ref.py)."""\n')
        with open(os.path.join(pkg, "preprocessing.py"), "w") as fh:
            fh.write('"""Empty stub: sequence.pyx imports deepgrp.preprocessing but never uses it."""\n')
        return True
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


def load():
    """Import the compiled reference modules; returns (sequence, mss) or None if unavailable."""
    if not built():
        return None
    import importlib.util
    pkg = os.path.join(OUT, "deepgrp")
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    saved = {k: sys.modules.get(k) for k in ("deepgrp", "deepgrp.preprocessing",
                                              "deepgrp.sequence", "deepgrp.mss")}
    try:
        import types
        fake = types.ModuleType("deepgrp")
        fake.__path__ = [pkg]
        sys.modules["deepgrp"] = fake
        sys.modules["deepgrp.preprocessing"] = types.ModuleType("deepgrp.preprocessing")
        mods = []
        for mod in ("sequence", "mss"):
            spec = importlib.util.spec_from_file_location("deepgrp." + mod,
                                                          os.path.join(pkg, mod + suffix))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            mods.append(m)
        return tuple(mods)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "oracle/_ref NOT built (reference absent)")
