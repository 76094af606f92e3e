"""TEST INFRASTRUCTURE ONLY — stage the unmodified reference into oracle/_ref/ (git-ignored, but shipped to the
GPU box with the repository snapshot) so that bench.py's cpu_baseline can time the REAL reference class on the
box's host cores and the GPU tests can run the reference's own run.py against this repository's drop-in modules.

The reference is pure Python: there is nothing to compile; this recipe copies its files where they lie under
/root/reference.  It only runs where /root/reference exists (the build container); nothing here is imported by
the product package, and the copies never enter the git history.

    python oracle/make_ref.py
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MPPI_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("control.py", "sys_params.py", "utils.py", "run.py", "xydq_circle.txt", "xydq.txt", "trajectory.txt",
         "trajectory1.txt")


def stage() -> str | None:
    if not os.path.isfile(os.path.join(SRC, "control.py")):
        return None
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return DST


def staged_dir() -> str | None:
    """Where an unmodified reference can be imported from: the staged copy, else the checkout itself."""
    for d in (DST, SRC):
        if os.path.isfile(os.path.join(d, "control.py")):
            return d
    return None


if __name__ == "__main__":
    print(stage() or f"{SRC} not found: nothing staged")
