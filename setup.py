"""Packaging entry point (reference: /root/reference/setup.py:1-10 is a bare ``find_packages()``).

``python setup.py build_ext --inplace`` (or ``pip install -e .``) compiles ``csrc/*.cu`` for sm_100a into the
in-tree C-ABI library ``multimodalreactiongeneration_b200/csrc/libmrg_b200.so`` through the same recipe as
``__graft_entry__.build()`` (``multimodalreactiongeneration_b200/_build.py``: plain nvcc, no torch extension machinery —
the library has no torch types in its ABI)."""
import os
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class build_ext(Command):
    description = "compile csrc/*.cu into libmrg_b200.so (sm_100a)"
    user_options = [("inplace", "i", "accepted for compatibility; the library is always built in-tree"),
                    ("force", "f", "rebuild even if the library is newer than its sources")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace = 0
        self.force = 0

    def finalize_options(self):
        pass

    def run(self):
        from multimodalreactiongeneration_b200 import _build
        print(_build.build(force=bool(self.force)))


class build_py_with_lib(build_py):
    def run(self):
        self.run_command("build_ext")
        super().run()


setup(
    name="multimodalreactiongeneration_b200",
    version="0.2.0",
    description="B200-native (sm_100a) LSTM hot path of MultimodalReactionGeneration behind the reference's module API",
    packages=find_packages(include=["multimodalreactiongeneration_b200", "multimodalreactiongeneration_b200.*"]),
    package_data={"multimodalreactiongeneration_b200": ["csrc/*.so", "csrc/*.cu", "csrc/*.cuh"]},
    cmdclass={"build_ext": build_ext, "build_py": build_py_with_lib},
)
