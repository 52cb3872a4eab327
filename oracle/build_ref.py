"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that builds the reference itself for the CPU arm.

The reference is Python, so "building" it means byte-compiling its own sources, where they lie under
/root/reference, into ``oracle/_ref/`` (marshalled code objects, ``*.pycode``, git-ignored, NOT gpurun-ignored: they travel
to the GPU box like a compiled ``.so`` would; the extension is not ``.pyc`` because snapshot tools tend to drop those).  No reference source is copied into the repository.  The box runs
the same image (same CPython), so the byte code loads there; ``ref_harness.reference_modules(compiled=True)``
imports it with TensorFlow / Keras / matplotlib stubbed exactly as for the source tree.

    python oracle/build_ref.py            # no-op when /root/reference is not mounted
"""
import importlib.abc
import importlib.machinery
import importlib.util
import marshal
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_DIR = os.environ.get("CK_REFERENCE_DIR", "/root/reference")
# the modules the self-play / arena path imports (training_pipeline.py:31-40, MCTS.py, Checkers.py)
FILES = ["Checkers.py", "MCTS.py", "training_pipeline.py", "CLR/__init__.py", "CLR/clr_callback.py", "LRFinder/keras_callback.py"]


def build(force=False):
    """-> output directory, or None when the reference is not available here"""
    if not os.path.isfile(os.path.join(REFERENCE_DIR, "Checkers.py")):
        return OUT if os.path.isfile(os.path.join(OUT, "Checkers.pycode")) else None
    for rel in FILES:
        src = os.path.join(REFERENCE_DIR, rel)
        dst = os.path.join(OUT, rel[:-3] + ".pycode")
        if not os.path.isfile(src):
            continue
        if not force and os.path.isfile(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, "rb") as f:
            code = compile(f.read(), rel, "exec", dont_inherit=True)
        with open(dst, "wb") as f:
            marshal.dump(code, f)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d\n" % sys.version_info[:2])
    return OUT


def available():
    try:
        ver = open(os.path.join(OUT, "PYTHON_VERSION")).read().strip()
    except OSError:
        return False
    return ver == "%d.%d" % sys.version_info[:2] and os.path.isfile(os.path.join(OUT, "Checkers.pycode"))


class _Loader(importlib.abc.Loader):
    def __init__(self, path):
        self.path = path

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(self.path, "rb") as f:
            code = marshal.load(f)
        module.__file__ = self.path
        exec(code, module.__dict__)


class Finder(importlib.abc.MetaPathFinder):
    """meta-path finder for the compiled reference modules in oracle/_ref (installed by ref_harness for the duration of
    an import of the reference, removed afterwards)"""

    def find_spec(self, fullname, path=None, target=None):
        rel = fullname.replace(".", os.sep)
        f = os.path.join(OUT, rel + ".pycode")
        if os.path.isfile(f):
            return importlib.util.spec_from_loader(fullname, _Loader(f), origin=f)
        pkg = os.path.join(OUT, rel)
        if os.path.isdir(pkg):
            init = os.path.join(pkg, "__init__.pycode")
            loader = _Loader(init) if os.path.isfile(init) else None
            spec = importlib.machinery.ModuleSpec(fullname, loader, origin=init if loader else None, is_package=True)
            spec.submodule_search_locations = [pkg]
            return spec
        return None


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
