"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/conesgpu.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from cones_perception_b200 import api
from cones_perception_b200.params import PRESETS, GroundParams, load_yaml_params
from cones_perception_b200.pointcloud2 import PointCloud2, PointField, make_view

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def header_symbols():
    text = open(os.path.join(ROOT, "include", "conesgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(api.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/conesgpu.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == syms, "api.ABI_SYMBOLS out of sync with the header"


def test_abi_version_and_strerror():
    lib = api.load_library()
    assert lib.cp_abi_version() == 2
    assert lib.cp_strerror(0) == b"ok"
    assert b"capacity" in lib.cp_strerror(api.CP_E_CAPACITY)


def test_struct_layouts_match_header():
    from cones_perception_b200.params import CDetectParams, CGroundParams
    from cones_perception_b200.pointcloud2 import CCloudView
    assert ctypes.sizeof(CDetectParams) == 7 * 8 + 2 * 4 + 2 * 4
    assert ctypes.sizeof(CGroundParams) == 8
    assert ctypes.sizeof(CCloudView) == 48 and CCloudView.off_x.offset == 24
    assert ctypes.sizeof(api.CConfig) == 40
    assert api.CLUSTER_DTYPE.itemsize == 16 and api.COUNTER_DTYPE.itemsize == 32


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(api.ConesGpuError) as e:
        api.ConesGpu(max_points=1024)
    assert e.value.status == api.CP_E_CUDA
    assert "no CPU fallback" in e.value.detail


def test_create_rejects_bad_config():
    lib = api.load_library()
    h = ctypes.c_void_p()
    cfg = api.CConfig(0, 0, 1, 16, 0, 0)
    assert lib.cp_create(ctypes.byref(h), ctypes.byref(cfg)) == api.CP_E_PARAM
    assert lib.cp_create(None, ctypes.byref(cfg)) == api.CP_E_PARAM


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(FileNotFoundError):
        api.load_library(str(tmp_path / "libconesgpu.so"))


def test_field_resolution_like_fromROSMsg():
    a = np.zeros((4, 6), np.float32)
    msg = PointCloud2(data=a.view(np.uint8).reshape(-1), width=4, point_step=24,
                      fields=[PointField("x", 4), PointField("y", 8), PointField("z", 12),
                              PointField("ring", 16, datatype=4), PointField("intensity", 20)])
    v = make_view(msg, fake_missing_intensity=True)
    assert (v.off_x, v.off_y, v.off_z, v.off_intensity) == (4, 8, 12, 20)
    msg.fields = msg.fields[:4]
    assert make_view(msg, fake_missing_intensity=True).off_intensity == 0     # src/cone_detection.cpp:142-151
    assert make_view(msg, fake_missing_intensity=False).off_intensity == -1   # ground node: intensity = 0
    msg.fields[0] = PointField("x", 4, datatype=8)                            # FLOAT64 x does not match
    assert make_view(msg, fake_missing_intensity=True).off_x == -1


def test_yaml_presets_match_reference_files(tmp_path):
    text = {"our": (7.0, 0.7, -0.5, 90.0, 3, 50), "fsai": (6.0, 1.0, -0.09, 160.0, 3, 500),
            "simulation": (10.0, 1.0, -5.0, 160.0, 2, 500)}
    for name, (dmax, dmin, lvl, ang, mn, mx) in text.items():
        p = PRESETS[name]
        assert (p.distance_treshold_max, p.distance_treshold_min, p.level_threshold, p.angle_threshold,
                p.min_cluster_size, p.max_cluster_size) == (dmax, dmin, lvl, ang, mn, mx)
        assert p.voxel_filter_leaf_size_x == p.voxel_filter_leaf_size_y == p.voxel_filter_leaf_size_z == 0.04
    y = tmp_path / "p.yaml"
    y.write_text("distance_treshold_max: 6.0\nmax_cluster_size: 500\nnum_of_sectors: 16\n")
    d = load_yaml_params(str(y), PRESETS["our"])
    assert d.distance_treshold_max == 6.0 and d.max_cluster_size == 500 and d.distance_treshold_min == 0.7
    y.write_text("num_of_sectors: 16\ndefault_lowest_point: -0.1\n")
    g = load_yaml_params(str(y))
    assert isinstance(g, GroundParams) and g.default_lowest_point == -0.1
