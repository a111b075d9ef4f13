"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the UNMODIFIED reference (capaulson/pyKriging at /root/reference) inside this
build container so that (a) the NumPy restatement in ``oracle/oracle_np.py`` can be
pinned against it and (b) golden vectors can be generated (``oracle/make_golden.py``).

The reference does not import as-is here:
  * ``matplotlib``, ``pylab``, ``mpl_toolkits``, ``inspyred`` are not installed
    (``pyKriging/krige.py:9-17``)  -> empty stub modules are registered;
  * ``np.mat`` was removed in NumPy 2.0 (``pyKriging/matrixops.py:30``)
    -> ``numpy.mat = numpy.asmatrix`` shim.
No reference source is modified.  /root/reference does not exist on the GPU box; the committed recipe
``oracle/make_ref.py`` copies the reference's package, unmodified, into the git-ignored ``oracle/_ref/``
(which travels with the gpurun snapshot), and this loader falls back to that copy.  Nothing here ever reads
/root/reference at run time on the GPU box.

Two ways to load it:
  * ``load_reference()``: the reference exactly as it is (its own NumPy ``matrixops``) -- the oracle's anchor, the
    CPU baseline / reference arm of ``bench.py``;
  * ``load_reference_on_b200_mixin()``: INTEGRATION.md "Level A" -- the reference's UNMODIFIED ``krige.py`` /
    ``regressionkrige.py`` with ``pyKriging.matrixops`` replaced by ``pykriging_b200.matrixops`` (the one-line change
    of ``krige.py:7``, done through ``sys.modules`` so that no reference file is edited).  One per process.
"""
import os
import pickle
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve_root():
    env = os.environ.get("PYKRIGING_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if cand and os.path.isfile(os.path.join(cand, "pyKriging", "matrixops.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _resolve_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pyKriging", "matrixops.py"))


def _install_stubs():
    import numpy as np

    if not hasattr(np, "mat"):
        np.mat = np.asmatrix
    names = [
        "matplotlib", "matplotlib.pyplot", "pylab", "mpl_toolkits",
        "mpl_toolkits.mplot3d", "inspyred", "inspyred.ec", "inspyred.swarm",
    ]
    for name in names:
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []  # behave like a package for "from x import y"
            sys.modules[name] = mod
    # attributes that the reference's import statements pull out of the stubs
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits.mplot3d"].axes3d = types.ModuleType("axes3d")
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["inspyred"].ec = sys.modules["inspyred.ec"]
    sys.modules["inspyred"].swarm = sys.modules["inspyred.swarm"]


def load_reference():
    """Return (kriging, regression_kriging, testfunctions) classes of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from pyKriging.krige import kriging
    from pyKriging.regressionkrige import regression_kriging
    from pyKriging.testfunctions import testfunctions

    return kriging, regression_kriging, testfunctions


def load_reference_on_b200_mixin():
    """INTEGRATION.md Level A: (kriging, regression_kriging) of the unmodified reference, inheriting from
    ``pykriging_b200.matrixops.matrixops`` instead of the reference's NumPy mixin.  Must run in a process that has
    not imported the plain reference (``pyKriging.matrixops`` is bound once per process)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "pyKriging.krige" in sys.modules or "pyKriging.matrixops" in sys.modules:
        raise RuntimeError("the plain reference is already imported in this process")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pykriging_b200.matrixops as b200_matrixops

    sys.modules["pyKriging.matrixops"] = b200_matrixops  # what `from .matrixops import matrixops` (krige.py:7) finds
    from pyKriging.krige import kriging
    from pyKriging.regressionkrige import regression_kriging

    assert kriging.__mro__[1] is b200_matrixops.matrixops
    return kriging, regression_kriging


def load_sampling_plan(k, n):
    """The reference's cached optimal LHCs (python-2 protocol-0 pickles)."""
    path = os.path.join(REFERENCE_ROOT, "pyKriging", "sampling_plans", "lhc_%d_%d.pkl" % (k, n))
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")
