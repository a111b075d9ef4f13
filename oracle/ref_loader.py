"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the UNMODIFIED reference (capaulson/pyKriging at /root/reference) inside this
build container so that (a) the NumPy restatement in ``oracle/oracle_np.py`` can be
pinned against it and (b) golden vectors can be generated (``oracle/make_golden.py``).

The reference does not import as-is here:
  * ``matplotlib``, ``pylab``, ``mpl_toolkits``, ``inspyred`` are not installed
    (``pyKriging/krige.py:9-17``)  -> empty stub modules are registered;
  * ``np.mat`` was removed in NumPy 2.0 (``pyKriging/matrixops.py:30``)
    -> ``numpy.mat = numpy.asmatrix`` shim.
No reference source is modified or copied.  /root/reference does not exist on the GPU
box, so nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.
"""
import os
import pickle
import sys
import types

REFERENCE_ROOT = os.environ.get("PYKRIGING_REFERENCE_ROOT", "/root/reference")


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


def load_sampling_plan(k, n):
    """The reference's cached optimal LHCs (python-2 protocol-0 pickles)."""
    path = os.path.join(REFERENCE_ROOT, "pyKriging", "sampling_plans", "lhc_%d_%d.pkl" % (k, n))
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")
