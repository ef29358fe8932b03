"""Regenerates the committed fixtures under tests/golden/.  Run in the build container:

    python tests/golden/make_golden.py

* cone_crops.npz  — the 577 hand-labelled real cone crops shipped with the reference
  (/root/reference/cones_clouds/cones.pkl, written by the data-collection branch of
  scripts/color_classifier_server.py:95-104).  Imported here because /root/reference does
  not exist on the GPU box.  Only x, y, z, intensity and the per-cone lengths are kept.
* cfg*_golden.npz — outputs of the CPU oracle (canonical mode) on seeded synthetic scans.
  The reference has no golden vectors of its own ("parity unpinned"); these pin OUR oracle
  against regressions and give the GPU tests a second, committed target.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cones_perception_b200 import scans  # noqa: E402
from oracle import oracle as O  # noqa: E402


def cone_crops():
    import pandas as pd
    src = "/root/reference/cones_clouds/cones.pkl"
    df = pd.read_pickle(src)
    xs, lens, colors = [], [], []
    for _, r in df.iterrows():
        x, y, z, it = (np.asarray(r[c], np.float32) for c in ("x", "y", "z", "intensity"))
        xs.append(np.stack([x, y, z, it], 1))
        lens.append(len(x))
        try:
            colors.append(int(r["color"]))
        except (TypeError, ValueError):  # a few hand-typed labels carry stray terminal escapes
            colors.append(-1)
    pts = np.concatenate(xs).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cone_crops.npz"), points=pts, lengths=np.array(lens, np.int32),
                        colors=np.array(colors, np.int8))
    print("cone_crops:", len(lens), "cones,", len(pts), "points")


def synthetic(idx, seed):
    cfg = scans.config(idx)
    frame = scans.generate_config5(1, seed)[0] if idx == 5 else scans.generate(cfg, 1, seed)[0]
    cl, ctr, _ = O.detect(O.view_of_xyzi(frame), cfg.detect, cfg.ground, O.CANONICAL)
    digest = hashlib.sha256(frame.tobytes()).hexdigest()
    np.savez_compressed(os.path.join(HERE, f"cfg{idx}_seed{seed}_golden.npz"), clusters=cl,
                        counters=np.array([ctr.n_points, ctr.n_ground_kept, ctr.n_cropped, ctr.n_voxels,
                                           ctr.n_components, ctr.n_clusters, ctr.key_bits], np.int64),
                        input_sha256=np.array(digest))
    print(f"cfg{idx} seed {seed}: K={len(cl)} V={ctr.n_voxels} C={ctr.n_cropped} sha={digest[:12]}")


if __name__ == "__main__":
    cone_crops()
    for idx, seed in ((1, 0), (2, 0), (2, 7), (4, 0), (5, 0)):
        synthetic(idx, seed)
