#!/usr/bin/env python
"""Runs pcl_pin (built against the REAL PCL, see CMakeLists.txt) on the inputs written by
`python tests/golden/make_golden.py pcl_inputs <dir>` and packs its output into tests/golden/pcl_<name>.npz,
the files tests/test_pcl_pin.py consumes.  Needs only numpy — no oracle, no CUDA.

    python tests/golden/make_golden.py pcl_inputs /tmp/pcl_in          # any box that has this repo built
    cmake -S tools/pcl_pin -B /tmp/pcl_pin_build -DCMAKE_BUILD_TYPE=Release && cmake --build /tmp/pcl_pin_build
    python tools/pcl_pin/run.py --bin /tmp/pcl_pin_build/pcl_pin --inputs /tmp/pcl_in [--pcl-version 1.10.0]

Each npz holds: voxels float32 [V,4] (VoxelGrid's output order), labels int32 [V] (smallest voxel index of the
kept cluster a voxel belongs to, -1 if filtered), order int32 [K,2] ((min index, size) in extract()'s order),
input_sha256 (of the float32 input cloud), pcl_version, shim (1 when produced by the stand-in build: pins nothing).
"""
import argparse
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bin", required=True, help="the pcl_pin executable")
    ap.add_argument("--inputs", required=True, help="directory written by make_golden.py pcl_inputs")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--pcl-version", default="unknown")
    ap.add_argument("--shim", action="store_true", help="mark the goldens as coming from the stand-in build")
    args = ap.parse_args()
    manifest = json.load(open(os.path.join(args.inputs, "manifest.json")))
    with tempfile.TemporaryDirectory() as tmp:
        for m in manifest:
            leaf = [repr(float(v)) for v in m["leaf"]]
            subprocess.run([args.bin, m["name"], os.path.join(args.inputs, m["name"] + ".bin"), *leaf,
                            str(m["min_cluster_size"]), str(m["max_cluster_size"]), tmp], check=True)
            vox = np.load(os.path.join(tmp, m["name"] + "_voxels.npy"))
            lab = np.load(os.path.join(tmp, m["name"] + "_labels.npy"))
            order = np.load(os.path.join(tmp, m["name"] + "_order.npy")).reshape(-1, 2)
            np.savez_compressed(os.path.join(args.out, f"pcl_{m['name']}.npz"), voxels=vox, labels=lab, order=order,
                                input_sha256=np.array(m["sha256"]), pcl_version=np.array(args.pcl_version),
                                shim=np.array(1 if args.shim else 0))
            print(f"pcl_{m['name']}.npz: V={len(vox)} K={len(order)}")


if __name__ == "__main__":
    main()
