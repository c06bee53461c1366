"""`pip install .` builds bess_b200/libbess_b200.so with nvcc for sm_100a (bess_b200/build.py: one nvcc call per source,
cross-compiles without a GPU) and installs the Python front-end next to it -- the counterpart of the reference's
python/setup.py:1-72, which builds the SWIG module `_cbess` from src/*.cpp.  Needs nvcc (CUDA >= 12.8) on PATH or $NVCC."""
import importlib.util
import os
import shutil

from setuptools import setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


class BuildWithNvcc(build_py):
    def run(self):
        spec = importlib.util.spec_from_file_location("_bess_b200_build", os.path.join(HERE, "bess_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)  # the build script alone: importing the package would try to load the library
        mod.build(force=False, verbose=True)
        # the public headers travel with the package (bess_b200/include) so that C / C++ clients of an installed copy find them
        inc = os.path.join(HERE, "bess_b200", "include")
        os.makedirs(inc, exist_ok=True)
        for f in os.listdir(os.path.join(HERE, "include")):
            shutil.copy2(os.path.join(HERE, "include", f), os.path.join(inc, f))
        super().run()


setup(cmdclass={"build_py": BuildWithNvcc})
