"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference Python.

Puts ``/root/reference`` on ``sys.path`` with ``tensorflow`` / ``keras`` /
``matplotlib`` stubbed in ``sys.modules`` (none are installed here, there is no
network) so that ``Checkers.py``, ``MCTS.py`` and ``training_pipeline.py`` run
verbatim.  Used (a) by ``tests/golden/make_golden.py`` to produce the committed
golden vectors and (b) by the ``needs_reference`` CPU tests that pin the C
oracle (``oracle/ck_oracle.c``) against the real reference where it is mounted.

``/root/reference`` does not exist on the GPU box: nothing under ``-m gpu``,
``smoke()`` or ``bench.py`` imports this module.
"""
import contextlib
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_DIR = os.environ.get("CK_REFERENCE_DIR", "/root/reference")
COMPILED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # built by oracle/build_ref.py


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "Checkers.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    # a spec keeps importlib.util.find_spec(name) (PyTorch probes for TensorFlow that way) from raising
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_LOAD_MODEL_HOOK = {"fn": None}


def set_load_model(fn):
    """``fn(path) -> object with Keras-like .predict(x[B,8,8,14]) -> [p[B,512], v[B,1]]``."""
    _LOAD_MODEL_HOOK["fn"] = fn


def _load_model(path, *a, **k):
    if _LOAD_MODEL_HOOK["fn"] is None:
        raise RuntimeError("ref_harness.set_load_model() was not called")
    return _LOAD_MODEL_HOOK["fn"](path)


def install_stubs():
    """Stub list from SURVEY.md appendix A (training_pipeline.py:31-40 imports)."""
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "_ck_stub", False):
        return

    class Sequence(object):
        pass

    class Callback(object):
        def __init__(self, *a, **k):
            pass

    tf = _stub("tensorflow", _ck_stub=True)
    keras = _stub("tensorflow.keras")
    tf.keras = keras
    keras.utils = _stub("tensorflow.keras.utils", Sequence=Sequence)
    keras.callbacks = _stub("tensorflow.keras.callbacks", Callback=Callback)
    keras.backend = _stub("tensorflow.keras.backend")
    keras.models = _stub("tensorflow.keras.models", load_model=_load_model)
    k2 = _stub("keras")
    k2.callbacks = _stub("keras.callbacks", Callback=Callback)
    k2.backend = _stub("keras.backend")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")


_REF_MODULE_NAMES = ("Checkers", "MCTS", "training_pipeline", "TicTacToe",
                     "CLR", "CLR.clr_callback", "LRFinder", "LRFinder.keras_callback")


def compiled_reference_available():
    """the reference's own byte code under oracle/_ref (oracle/build_ref.py): what travels to the GPU box"""
    from . import build_ref
    return build_ref.available()


@contextlib.contextmanager
def reference_modules(with_pipeline=False, compiled=False):
    """Context manager yielding a namespace with the reference's own modules.
    ``compiled``: import the byte-compiled copy in oracle/_ref instead of the source tree (bench.py's CPU arm).

    The repo ships drop-in modules with the same names (``Checkers``, ``MCTS``,
    ``training_pipeline``); to keep the two apart the reference copies are
    imported under a temporary ``sys.path``/``sys.modules`` and removed again on
    exit, and the caller keeps the module objects.
    """
    ref_dir = COMPILED_DIR if compiled else REFERENCE_DIR
    if compiled and not compiled_reference_available():
        raise RuntimeError("compiled reference not built (python oracle/build_ref.py where /root/reference is mounted)")
    if not compiled and not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_DIR)
    saved = {n: sys.modules.pop(n) for n in _REF_MODULE_NAMES if n in sys.modules}
    saved_path = list(sys.path)
    finder = None
    if compiled:
        from . import build_ref
        finder = build_ref.Finder()
        sys.meta_path.insert(0, finder)
    else:
        sys.path.insert(0, ref_dir)
    try:
        ns = types.SimpleNamespace()
        ns.Checkers = importlib.import_module("Checkers")
        ns.MCTS = importlib.import_module("MCTS")
        if with_pipeline:
            install_stubs()
            ns.training_pipeline = importlib.import_module("training_pipeline")
        assert os.path.dirname(ns.Checkers.__file__) == ref_dir
        yield ns
    finally:
        sys.path[:] = saved_path
        if finder is not None:
            sys.meta_path.remove(finder)
        for n in _REF_MODULE_NAMES:
            sys.modules.pop(n, None)
        sys.modules.update(saved)


def load_reference(with_pipeline=False):
    """Non-context variant: returns the namespace, leaves sys.modules clean."""
    with reference_modules(with_pipeline) as ns:
        return ns
