"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the snapshot).

    python tools/install_reference.py [--force]

The GPU box has no /root/reference, so the tests that drive the reference's own modules through the B200 path
(tests/test_reference_gpu.py) need an importable copy that travels.  This is an install, not a vendoring: nothing
under baseline/_ref is tracked by git.

Step 1 is the contract's pip line (from a /tmp copy because the build writes egg-info into the source tree):
    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>
The reference's setup.py uses `find_packages()`, which silently skips every directory without an __init__.py --
reconstruction/vision, reconstruction/autoencoder, all of policies/ -- and ships no package data, so the chart
meshes the models load at start-up (objects/vision_charts.obj, touch_chart.obj, test_objects/*.obj) are missing too.
Step 2 completes the install with exactly those files (the .py modules pip skipped + the .obj assets), byte for byte.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PTK_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
ASSETS = ("objects/vision_charts.obj", "objects/touch_chart.obj", "objects/data_split.npy", "objects/test_objects/0.obj",
          "objects/test_objects/1.obj")


def install(force=False):
    marker = os.path.join(DST, "pterotactyl", "reconstruction", "vision", "model.py")
    if os.path.exists(marker) and not force:
        return DST
    if not os.path.isdir(os.path.join(SRC, "pterotactyl")):
        raise FileNotFoundError(f"{SRC} holds no pterotactyl checkout (the GPU box only uses the prebuilt install)")
    tmp = "/tmp/ptk_reference_src"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(SRC, tmp, ignore=shutil.ignore_patterns("images", "notebook", "*.stl", "*.STL", "*.ttf"))
    if force:
        shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade",
           "--find-links", "/opt/wheelhouse", "--target", DST, tmp]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("pip install of the reference failed")
    # step 2: what find_packages() skipped
    n = 0
    for base, _dirs, files in os.walk(os.path.join(SRC, "pterotactyl")):
        rel = os.path.relpath(base, SRC)
        for f in files:
            if f.endswith(".py") or os.path.join(os.path.relpath(base, os.path.join(SRC, "pterotactyl")), f) in ASSETS:
                out = os.path.join(DST, rel, f)
                if not os.path.exists(out):
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    shutil.copyfile(os.path.join(base, f), out)
                    n += 1
    shutil.rmtree(tmp, ignore_errors=True)
    print(f"reference installed into {DST} (+{n} files pip's find_packages() skipped)")
    return DST


if __name__ == "__main__":
    install(force="--force" in sys.argv)
