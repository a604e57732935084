"""TEST INFRASTRUCTURE — not product code.

Stages the UNMODIFIED reference scripts where they can travel to the GPU box: copies the five Python scripts of
/root/reference (+ LICENSE.md) byte for byte into the git-ignored directory oracle/_ref/ .  `oracle/_ref/` is listed in
.gitignore (no reference source ever enters the history) but not in .gpurunignore, so it ships with the snapshot like a
built .so.  `bench.py --impl reference` then times the reference's OWN forward (`TEST1.RelightNet.forward`, imported
through oracle/ref_shims.py: cpu_baseline.kind = "reference"); without oracle/_ref it falls back to the oracle port
(kind = "port").  The weights are not copied: tests/golden/model_epoch99.pth is byte-identical to
/root/reference/model/model_epoch99.pth (md5 checked below).

    python oracle/fetch_ref.py            # called by __graft_entry__.build() when /root/reference exists
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("GFR_REFERENCE_SRC", "/root/reference")
FILES = (
    "train_raytracing_relighting_CelebAHQ_DSSIM_8x.py", "test_relight_single_image.py",
    "test_raytracing_relighting_CelebAHQ_DSSIM_8x.py", "test_relight_single_image_lighting_transfer.py",
    "train_lighting_transfer.py", "LICENSE.md",
)
WEIGHTS = (("model/model_epoch99.pth", "model_epoch99.pth"), ("model_lighting_transfer/model_epoch106.pth", "model_epoch106.pth"))


def _md5(p):
    h = hashlib.md5()
    with open(p, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def fetch(verbose=False):
    """-> True if oracle/_ref holds the reference scripts afterwards."""
    if not os.path.isfile(os.path.join(SRC, FILES[1])):
        return os.path.isfile(os.path.join(DEST, FILES[1]))
    os.makedirs(DEST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DEST, f))
    golden = os.path.join(HERE, "..", "tests", "golden")
    with open(os.path.join(DEST, "MANIFEST.txt"), "w") as m:
        m.write("byte copies of %s (made by oracle/fetch_ref.py; git-ignored)\n" % SRC)
        for f in FILES:
            m.write("%s  %s\n" % (_md5(os.path.join(DEST, f)), f))
        for rel, name in WEIGHTS:
            a, b = _md5(os.path.join(SRC, rel)), _md5(os.path.join(golden, name))
            if a != b:
                raise RuntimeError("tests/golden/%s is not the reference's %s" % (name, rel))
            m.write("%s  %s == tests/golden/%s\n" % (a, rel, name))
    if verbose:
        print("oracle/_ref: %d files" % len(FILES))
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch(verbose=True) else 1)
