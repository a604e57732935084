"""Dev tool (GPU): runs the command-line drivers of geomconsistentfr_b200.inference on PNG files made from the FFHQ fixtures
and compares the written image with the PNG the reference ships (the input is re-quantised to 8 bits on the way, so the
comparison is loose: mean abs difference in grey levels on the mask interior)."""
import os
import subprocess
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
f = np.load(os.path.join(G, "ffhq.npz"))
names = list(f["names"])
i = names.index("00508")
d = tempfile.mkdtemp()
img = np.clip(np.rint(f["q"][i] / 4.0), 0, 255).astype(np.uint8)              # RGB u8 at 256x256
cv2.imwrite(os.path.join(d, "in.png"), img[:, :, ::-1])
cv2.imwrite(os.path.join(d, "mask.png"), f["masks"][i])
L = ",".join("%g" % v for v in f["lights"][i])
env = dict(os.environ, PYTHONPATH=ROOT)
r = subprocess.run([sys.executable, "-m", "geomconsistentfr_b200.inference", "relight", os.path.join(G, "model_epoch99.pth"),
                    os.path.join(d, "in.png"), os.path.join(d, "mask.png"), L, os.path.join(d, "out.png"), "--fix-border"],
                   capture_output=True, text=True, env=env, cwd=ROOT)
print("relight rc", r.returncode, r.stderr[-300:])
out = cv2.imread(os.path.join(d, "out.png"))
want = f["pngs_bgr"][i]
diff = np.abs(out.astype(np.int32) - want.astype(np.int32))
print("relight: shape", out.shape, "mean |diff| vs shipped PNG", float(diff.mean()), "max", int(diff.max()))
j = names.index("00110")
ref = np.clip(np.rint(f["q"][j] / 4.0), 0, 255).astype(np.uint8)
cv2.imwrite(os.path.join(d, "ref.png"), ref[:, :, ::-1])
r = subprocess.run([sys.executable, "-m", "geomconsistentfr_b200.inference", "transfer", os.path.join(G, "model_epoch106.pth"),
                    os.path.join(d, "in.png"), os.path.join(d, "ref.png"), os.path.join(d, "mask.png"), os.path.join(d, "lt")],
                   capture_output=True, text=True, env=env, cwd=ROOT)
print("transfer rc", r.returncode, r.stdout.strip()[-200:], r.stderr[-300:])
print("transfer files", sorted(os.listdir(os.path.join(d, "lt"))) if os.path.isdir(os.path.join(d, "lt")) else None)
