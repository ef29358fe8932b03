"""Regenerates the committed fixtures under tests/golden/.  Run in the build container:

    python tests/golden/make_golden.py

* cone_crops.npz  — the 577 hand-labelled real cone crops shipped with the reference
  (/root/reference/cones_clouds/cones.pkl, written by the data-collection branch of
  scripts/color_classifier_server.py:95-104).  Imported here because /root/reference does
  not exist on the GPU box.  Only x, y, z, intensity and the per-cone lengths are kept.
* cone_images.npz — PINNED golden: the reference's own ColorClassifier.to_image
  (/root/reference/scripts/color_classifier_server.py:130-156, plain numpy/scipy, executed
  here from its source) on the 577 real crops and on seeded synthetic crops, including the
  inputs on which it raises (IndexError / interp1d ValueError -> flag, zero image).
* reference_nodes.npz — PINNED golden: outputs of the reference's own node sources compiled unmodified
  (oracle/_ref/libconesref.so via oracle/ref_shim, -O0 like the reference's catkin build): the ground node on
  config-2 scans (sha256 of the published cloud + survivors), the crop lambda's verdict for every point of the
  threshold-boundary clouds of the three presets, and the clouds the detection node publishes over 5-frame
  sequences (PCL's VoxelGrid / clustering inside it are the oracle's restatement).
* cfg*_golden.npz — outputs of the CPU oracle (canonical mode) on seeded synthetic scans.
  The reference has no golden vectors of its own; these pin OUR oracle against regressions (PCL stages
  included) and give the GPU tests a second, committed target.
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


def reference_to_image():
    """The reference's to_image, compiled from its own source (the module itself imports rospy and
    tensorflow, which are not installed, so only the constants and the function body are taken)."""
    import ast
    from scipy.interpolate import interp1d
    tree = ast.parse(open("/root/reference/scripts/color_classifier_server.py").read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.Assign):
            keep.append(node)
        if isinstance(node, ast.ClassDef):
            for m in node.body:
                if isinstance(m, ast.FunctionDef) and m.name == "to_image":
                    m.decorator_list = []
                    keep.append(m)
    ns = {"np": np, "interp1d": interp1d}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "color_classifier_server.py", "exec"), ns)
    return ns["to_image"]


def synthetic_crops(seed=0, n_crops=256):
    """Seeded cone-sized crops: box +-0.152 m around a centre 1.2..20 m away, any azimuth; a few
    degenerate ones (single point, repeated point, intensity / elevation outside the image)."""
    rng = np.random.default_rng(seed)
    crops = []
    for i in range(n_crops):
        r, az = rng.uniform(1.2, 20.0), rng.uniform(-np.pi, np.pi)
        cx, cy = r * np.cos(az), r * np.sin(az)
        n = int(rng.integers(1, 120))
        x = cx + rng.uniform(-0.152, 0.152, n)
        y = cy + rng.uniform(-0.152, 0.152, n)
        z = rng.uniform(-0.3, 0.3, n) if r > 2 else rng.uniform(-0.25, 0.25, n)
        it = rng.uniform(0, 255, n)
        kind = i % 16
        if kind == 11:
            n = 1
        elif kind == 12:
            x[:] = x[0]; y[:] = y[0]                     # one azimuth: zero horizontal range
        elif kind == 13:
            it[rng.integers(0, n)] = 255.5               # interp1d raises
        elif kind == 14:
            z[rng.integers(0, n)] = 0.9 * r              # elevation > 15 deg: IndexError
        elif kind == 15:
            z[rng.integers(0, n)] = -0.5 * r             # elevation < -15 deg: rows 0..14 still valid
        crops.append(np.stack([x, y, z, it], 1)[:n].astype(np.float32))
    return crops


def cone_images():
    to_image = reference_to_image()
    z = np.load(os.path.join(HERE, "cone_crops.npz"))
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    real = [z["points"][off[i]:off[i + 1]] for i in range(len(z["lengths"]))]
    synth = synthetic_crops()

    def run(crops):
        imgs, raised = np.zeros((len(crops), 15, 12), np.uint8), np.zeros(len(crops), np.uint8)
        for i, c in enumerate(crops):
            q = c.astype(np.float64)  # pc2.read_points yields Python floats
            row = {"x": list(q[:, 0]), "y": list(q[:, 1]), "z": list(q[:, 2]), "intensity": list(q[:, 3])}
            try:
                imgs[i] = to_image(row)[:, :, 0]
            except IndexError:
                raised[i] = 2
            except ValueError:
                raised[i] = 4
        return imgs, raised

    real_img, real_raised = run(real)
    syn_img, syn_raised = run(synth)
    np.savez_compressed(os.path.join(HERE, "cone_images.npz"), real_images=real_img, real_raised=real_raised,
                        synth_points=np.concatenate(synth), synth_lengths=np.array([len(c) for c in synth], np.int32),
                        synth_images=syn_img, synth_raised=syn_raised)
    print("cone_images:", len(real), "real (raised:", int((real_raised > 0).sum()), "),", len(synth),
          "synthetic (raised:", int((syn_raised > 0).sum()), ")")


def reference_nodes():
    """Runs the real reference nodes (oracle/_ref) and stores what they publish."""
    from cones_perception_b200.params import PRESETS
    from oracle import ref as R
    from tests.test_reference_pin import node_params, outside_sector_16
    from tests.util import boundary_cloud
    out = {}
    # ground_removal node: config 2, azimuths of the undefined 17th sector slot left out (Appendix C Q2)
    for seed in (0, 9):
        frame = outside_sector_16(scans.generate(scans.config(2), 1, base_seed=seed)[0])
        node = R.GroundNode()
        cloud32, meta = node.handle(frame)
        node.close()
        kept = int((cloud32[:, 3] == 1.0).sum() if False else np.count_nonzero(np.any(cloud32[:, :3] != 0, axis=1)))
        out[f"ground_seed{seed}_sha256"] = np.array(hashlib.sha256(cloud32.tobytes()).hexdigest())
        out[f"ground_seed{seed}_input_sha256"] = np.array(hashlib.sha256(frame.tobytes()).hexdigest())
        out[f"ground_seed{seed}_kept"] = np.array(kept)
        out[f"ground_seed{seed}_head"] = cloud32[:256].copy()
        out[f"ground_seed{seed}_meta"] = np.array(meta)
    # the crop lambda (src/cone_detection.cpp:191-203), one verdict per point, via single-point clouds
    for name in ("our", "fsai", "simulation"):
        d = PRESETS[name]
        cloud = boundary_cloud(d, seed=2, n_random=2000)
        cloud = np.ascontiguousarray(cloud[np.isfinite(cloud).all(1)])
        node = R.DetectNode(service=False, **node_params(d, classify_colors=False, use_points_buffer=False,
                                                        min_cluster_size=1, max_cluster_size=100000))
        # a point inside every preset's crop keeps prev_detected_cones non-empty, so a surviving point is published
        prime = np.array([[max(d.distance_treshold_min, 0.0) + 0.5, 0.0, max(d.level_threshold, -1.0) + 0.3, 1.0]], np.float32)
        assert len(node.handle(prime)[0]) == 0 and len(node.handle(prime)[0]) == 1
        keep = np.zeros(len(cloud), np.uint8)
        for i, p in enumerate(cloud):
            got = node.handle(p[None, :].copy())
            keep[i] = len(got[0]) == 1
            if not keep[i]:
                node.handle(prime)
        node.close()
        out[f"crop_{name}_cloud"] = cloud
        out[f"crop_{name}_keep"] = keep
    # cone_detection node over a sequence (config 1, `our` preset), both gate modes
    cfg = scans.config(1)
    frames = scans.generate(cfg, 5, base_seed=40)
    for buffer in (True, False):
        node = R.DetectNode(service=False, **node_params(cfg.detect, classify_colors=False, use_points_buffer=buffer))
        for fi, f in enumerate(frames):
            got = node.handle(f)
            out[f"detect_buffer{int(buffer)}_frame{fi}"] = got[0]
        node.close()
    np.savez_compressed(os.path.join(HERE, "reference_nodes.npz"), **out)
    print("reference_nodes:", {k: (v.shape if v.ndim else v.item()) for k, v in out.items() if "cloud" not in k and "head" not in k})


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


def dam_net():
    """dam_net_model.npz — the reference's classifier file (models/dam_net/dam_net.tflite, 23.5 kB) carried as a
    byte array, because /root/reference does not exist on the GPU box, with the numpy forward pass's outputs
    (oracle/dam_net_ref.py) on the 577 recorded crops' range images as a regression pin."""
    from cones_perception_b200 import tflite_model
    from oracle import dam_net_ref as D
    raw = open("/root/reference/models/dam_net/dam_net.tflite", "rb").read()
    g = tflite_model.load(raw)
    imgs = np.load(os.path.join(HERE, "cone_images.npz"))["real_images"]
    probs, logits = zip(*[D.forward(g, im) for im in imgs])
    # the same network as the Keras SavedModel the reference ships beside the .tflite file
    # (models/dam_net/variables/variables.data-00000-of-00001): the head of the tensor bundle holds, back to back,
    # conv2d/kernel (3,3,1,16 HWIO) conv2d/bias conv2d_1/kernel (3,3,16,32) conv2d_1/bias, batch_normalization
    # gamma / beta / moving_mean / moving_variance (32 each), dense/kernel (64,3), dense/bias — the optimizer
    # slots that follow are not needed
    kv = open("/root/reference/models/dam_net/variables/variables.data-00000-of-00001", "rb").read()[:20492]
    np.savez_compressed(os.path.join(HERE, "dam_net_model.npz"), tflite=np.frombuffer(raw, np.uint8),
                        keras_variables=np.frombuffer(kv, np.uint8),
                        sha256=np.array(hashlib.sha256(raw).hexdigest()),
                        real_probs=np.stack(probs).astype(np.float32), real_logits=np.stack(logits).astype(np.float32))
    print("dam_net:", len(raw), "bytes,", len(imgs), "images")


PCL_PIN_CASES = ((1, 0), (2, 0), (3, 5), (4, 0), (5, 0))      # (config, seed): one frame of each BASELINE.json config


def pcl_pin_input(idx, seed):
    """The cloud the reference hands to pcl::VoxelGrid for one frame of config idx: the crop survivors
    (ground removal + zero padding + filter_points_position), x, y, z, intensity as float32 [C,4]."""
    from tests.util import oracle_stages
    cfg = scans.config(idx)
    frame = scans.generate_config5(1, seed)[0] if idx == 5 else scans.generate(cfg, 1, seed)[0]
    st = oracle_stages(frame, cfg.detect, cfg.ground)
    c = st["cropped"]
    pts = np.ascontiguousarray(np.stack([c["x"], c["y"], c["z"], c["intensity"]], 1), np.float32)
    return cfg, pts


def pcl_inputs(out_dir):
    """Inputs of tools/pcl_pin (run where the real PCL exists): <out_dir>/<name>.bin + manifest.json."""
    import json
    os.makedirs(out_dir, exist_ok=True)
    manifest = []
    for idx, seed in PCL_PIN_CASES:
        cfg, pts = pcl_pin_input(idx, seed)
        name = f"cfg{idx}_seed{seed}"
        pts.tofile(os.path.join(out_dir, name + ".bin"))
        d = cfg.detect
        manifest.append({"name": name, "points": int(len(pts)), "sha256": hashlib.sha256(pts.tobytes()).hexdigest(),
                         "leaf": [d.voxel_filter_leaf_size_x, d.voxel_filter_leaf_size_y, d.voxel_filter_leaf_size_z],
                         "min_cluster_size": d.min_cluster_size, "max_cluster_size": d.max_cluster_size})
        print(f"pcl input {name}: {len(pts)} survivors")
    json.dump(manifest, open(os.path.join(out_dir, "manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "pcl_inputs":
        pcl_inputs(sys.argv[2])
        sys.exit(0)
    cone_crops()
    cone_images()
    dam_net()
    reference_nodes()
    for idx, seed in ((1, 0), (2, 0), (2, 7), (4, 0), (5, 0)):
        synthetic(idx, seed)
