"""Stage the UNMODIFIED reference sources of the hot path under baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that
the GPU box -- which has no /root/reference -- can run the reference's own model.py: bench.py's `--impl reference` arm,
the PyTorch-on-B200 denominator of north_star's >= 15x target, and the live-reference oracle checks.

Test / measurement infrastructure only.  Nothing under baseline/_ref/ is tracked by git or imported by the product
package; oracle.live_reference applies the single in-memory reporting shim (`.data[0]` -> `.item()`) at load.

    python oracle/stage_ref.py        # run in the build container (needs /root/reference)
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/augmented_cyclegan"
DST = os.path.join(ROOT, "baseline", "_ref", "augmented_cyclegan")
FILES = ("modules.py", "networks.py", "model.py")


def stage():
    """copy modules.py / networks.py / model.py byte for byte; returns True when the staged tree is complete"""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return all(os.path.isfile(os.path.join(DST, f)) for f in FILES)


if __name__ == "__main__":
    print("staged" if stage() else "reference not available", DST)
