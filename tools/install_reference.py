"""Install the UNMODIFIED reference tree into baseline/_ref (git-ignored; it travels to the GPU box with gpurun).

    python tools/install_reference.py [--src /root/reference]

The contract's recipe — `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference` — is tried first and recorded; TransEditor is a script tree without setup.py / pyproject.toml, so pip has
nothing to build and the fallback copies the source files byte for byte (Python sources + the two CUDA extension
sources, ~4 MB; datasets, FID statistics and result images are left out).  baseline/_ref is used ONLY as
  * the tree tools/run_ref_script.py runs the reference's own train / test scripts from (against this repository's
    modules), and
  * the GPU-side oracle / baseline: the reference's CUDA ops JIT-built for sm_100a and its own nn.Modules on cuDNN
    (tests/test_gpu_reference_*.py, tools/reference_gpu_baseline.py) — never by the product path.
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_DIRS = ("utils", "our_interfaceGAN", "pSp")
SKIP_EXT = (".pkl", ".npy", ".png", ".jpg", ".jpeg", ".pt", ".pth", ".gif", ".pyc")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    if not os.path.isfile(os.path.join(a.src, "model_spatial_query.py")):
        sys.exit("reference tree not found at %s" % a.src)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                          "--find-links", "/opt/wheelhouse", "--target", DST, a.src], capture_output=True, text=True)
    note = "pip install rc=%d: %s" % (pip.returncode, (pip.stderr.strip().splitlines() or ["ok"])[-1][:200])
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    n, digest = 0, hashlib.sha256()
    for name in sorted(os.listdir(a.src)):
        full = os.path.join(a.src, name)
        if os.path.isfile(full) and name.endswith((".py", ".md", ".yaml")) or name == "LICENSE":
            shutil.copy2(full, os.path.join(DST, name))
            digest.update(open(full, "rb").read())
            n += 1
    for d in KEEP_DIRS:
        for base, _, files in os.walk(os.path.join(a.src, d)):
            for f in sorted(files):
                if f.lower().endswith(SKIP_EXT):
                    continue
                src = os.path.join(base, f)
                rel = os.path.relpath(src, a.src)
                os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                shutil.copy2(src, os.path.join(DST, rel))
                digest.update(open(src, "rb").read())
                n += 1
    with open(os.path.join(DST, "INSTALL_NOTE.txt"), "w") as f:
        f.write("%s\ncopied %d files from %s (unmodified), sha256 of contents %s\n" % (note, n, a.src, digest.hexdigest()))
    print(note)
    print("copied %d files into %s" % (n, DST))


if __name__ == "__main__":
    main()
