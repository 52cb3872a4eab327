"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that builds the reference itself for the CPU arm.

The reference is Python, so "building" it means byte-compiling its own sources, where they lie under
/root/reference, into ``oracle/_ref/`` (sourceless ``.pyc`` files, git-ignored, NOT gpurun-ignored: they travel
to the GPU box like a compiled ``.so`` would).  No reference source is copied into the repository.  The box runs
the same image (same CPython), so the byte code loads there; ``ref_harness.reference_modules(compiled=True)``
imports it with TensorFlow / Keras / matplotlib stubbed exactly as for the source tree.

    python oracle/build_ref.py            # no-op when /root/reference is not mounted
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_DIR = os.environ.get("CK_REFERENCE_DIR", "/root/reference")
# the modules the self-play / arena path imports (training_pipeline.py:31-40, MCTS.py, Checkers.py)
FILES = ["Checkers.py", "MCTS.py", "training_pipeline.py", "CLR/__init__.py", "CLR/clr_callback.py", "LRFinder/keras_callback.py"]


def build(force=False):
    """-> output directory, or None when the reference is not available here"""
    if not os.path.isfile(os.path.join(REFERENCE_DIR, "Checkers.py")):
        return OUT if os.path.isfile(os.path.join(OUT, "Checkers.pyc")) else None
    for rel in FILES:
        src = os.path.join(REFERENCE_DIR, rel)
        dst = os.path.join(OUT, rel[:-3] + ".pyc")
        if not os.path.isfile(src):
            continue
        if not force and os.path.isfile(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # unchecked: the .pyc is valid without its source file next to it
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d\n" % sys.version_info[:2])
    return OUT


def available():
    try:
        ver = open(os.path.join(OUT, "PYTHON_VERSION")).read().strip()
    except OSError:
        return False
    return ver == "%d.%d" % sys.version_info[:2] and os.path.isfile(os.path.join(OUT, "Checkers.pyc"))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
