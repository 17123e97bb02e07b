"""Recipe: compile the reference's only native component (the legacy Cython models) into
``oracle/_ref/`` so the golden generator can use the REAL ``HyperSphere`` of the reference.

TEST INFRASTRUCTURE ONLY.  Sources are compiled from where they lie under
/root/reference (copied to a /tmp scratch package because cythonize writes next to the
.pyx and /root/reference is read-only); only the built extension lands in the repo, under
the git-ignored ``oracle/_ref/``.  Nothing here runs on the GPU box.

Needs the Cython directive ``cpow=True`` (Cython 3.x rejects ``ngrid**i_dim`` at
``harmonic/model_legacy.pyx:433`` otherwise).

The flow half of the reference (jax / flax / tfp / distrax) is NOT buildable here: those
packages are absent from the image and the wheelhouse, and there is no network.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HARMONIC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

SETUP = """
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
setup(
    name="harmonic_legacy_ref",
    ext_modules=cythonize(
        [Extension("harmonic.model_legacy", ["harmonic/model_legacy.pyx"],
                   include_dirs=[numpy.get_include()])],
        compiler_directives={"language_level": "2", "cpow": True},
    ),
)
"""


def build(force=False):
    if not os.path.isdir(os.path.join(REF, "harmonic")):
        return None
    pkg_out = os.path.join(OUT, "harmonic")
    if not force and os.path.isdir(pkg_out) and any(
        f.startswith("model_legacy") and f.endswith(".so") for f in os.listdir(pkg_out)
    ):
        return OUT
    tmp = tempfile.mkdtemp(prefix="harmonic_ref_build_")
    try:
        os.makedirs(os.path.join(tmp, "harmonic"))
        for f in ("model_legacy.pyx", "model_abstract.py", "logs.py", "default-logging-config.yaml"):
            src = os.path.join(REF, "harmonic", f)
            if os.path.exists(src):
                shutil.copy(src, os.path.join(tmp, "harmonic", f))
        open(os.path.join(tmp, "harmonic", "__init__.py"), "w").close()
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP)
        subprocess.check_call(
            [sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp
        )
        os.makedirs(pkg_out, exist_ok=True)
        for f in os.listdir(os.path.join(tmp, "harmonic")):
            if f.endswith(".so"):
                shutil.copy(os.path.join(tmp, "harmonic", f), os.path.join(pkg_out, f))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
