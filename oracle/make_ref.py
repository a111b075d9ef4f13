"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.

Recipe that ships the UNMODIFIED reference to the GPU box: copies the reference's Python package (and the two
cached sampling plans the golden cases use) from ``/root/reference`` into the git-ignored ``oracle/_ref/`` so that it
travels with the gpurun snapshot (``/root/reference`` does not exist there).  Nothing is edited; nothing under
``oracle/_ref/`` is tracked.

    python oracle/make_ref.py        # also run by __graft_entry__.build() whenever /root/reference is present

Consumers: ``oracle/ref_loader.py`` (falls back to ``oracle/_ref`` when ``/root/reference`` is absent) --
``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference"), the Level-A drop-in test
(``tests/test_gpu_level_a.py``) and ``tests/test_oracle.py``'s live comparison.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PYKRIGING_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["__init__.py", "krige.py", "regressionkrige.py", "matrixops.py", "samplingplan.py", "testfunctions.py",
         "utilities.py"]
PLANS = ["lhc_2_20.pkl", "lhc_2_25.pkl", "lhc_3_40.pkl"]


def make_ref(verbose=False):
    src_pkg = os.path.join(SRC, "pyKriging")
    if not os.path.isfile(os.path.join(src_pkg, "matrixops.py")):
        return False
    dst_pkg = os.path.join(DST, "pyKriging")
    os.makedirs(os.path.join(dst_pkg, "sampling_plans"), exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src_pkg, f), os.path.join(dst_pkg, f))
    for f in PLANS:
        shutil.copyfile(os.path.join(src_pkg, "sampling_plans", f), os.path.join(dst_pkg, "sampling_plans", f))
    for f in ("LICENSE", "LICENSE.txt", "LICENSE.md"):
        if os.path.isfile(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    if verbose:
        print("reference package copied to %s" % dst_pkg)
    return True


if __name__ == "__main__":
    ok = make_ref(verbose=True)
    sys.exit(0 if ok else 1)
