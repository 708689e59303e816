"""Recipe: stage the reference's own Python modules under oracle/_ref/ (git-ignored, travels with gpurun) so that the
GPU box can time the UNMODIFIED reference next to the GPU path (bench.py cpu_baseline.reference_python).

    python oracle/stage_ref.py            # in the build container, where /root/reference is mounted

Only snppipeline/*.py is taken (no data sets); nothing is edited.  TEST / BENCH INFRASTRUCTURE ONLY: the product
package never imports anything from here."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SNP_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage():
    src = os.path.join(SRC, "snppipeline")
    if not os.path.isfile(os.path.join(src, "pileup.py")):
        return False
    dst = os.path.join(DST, "snppipeline")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
    return True


def staged_root():
    """oracle/_ref when the reference's modules are staged there, else None."""
    return DST if os.path.isfile(os.path.join(DST, "snppipeline", "pileup.py")) else None


if __name__ == "__main__":
    print("staged" if stage() else "no reference at %s" % SRC)
    sys.exit(0)
