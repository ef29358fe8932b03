"""ros_shell/ is a catkin package (CMakeLists.txt + package.xml) that builds the drop-in nodes under the
reference's package and executable names.  ROS is not installed here, so the CMake logic is exercised against a
stand-in `catkin` package (tools/catkin_stub): configure, compile and link both node executables against the
stand-in ROS surface and the prebuilt libconesgpu, and check the install rules.  The CUDA branch (compiling
pipeline.cu for sm_100a inside CMake) is configured but not built here — build.py compiles the same file."""
import os
import shutil
import subprocess
import xml.etree.ElementTree as ET

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHELL = os.path.join(ROOT, "ros_shell")
STUB = os.path.join(ROOT, "tools", "catkin_stub")


def _fake_reference(tmp_path):
    """A directory shaped like a checkout of the reference (only names matter to the CMake logic)."""
    ref = tmp_path / "reference"
    for d in ("srv", "launch", "config", "scripts", "models/dam_net", "rviz"):
        (ref / d).mkdir(parents=True)
    (ref / "srv" / "ClassifyColorSrv.srv").write_text("sensor_msgs/PointCloud2[] cones_clouds\n---\nuint8[] colors\n")
    (ref / "scripts" / "color_classifier_server.py").write_text("#!/usr/bin/env python3\n")
    (ref / "launch" / "cones_detection.launch").write_text("<launch/>\n")
    (ref / "config" / "ground_removal_params.yaml").write_text("num_of_sectors: 16\n")
    return ref


def test_package_xml_keeps_the_reference_package_name_and_drops_pcl():
    root = ET.parse(os.path.join(SHELL, "package.xml")).getroot()
    assert root.find("name").text == "cones_perception"          # launch files say pkg="cones_perception"
    deps = {e.text for e in root.iter() if e.tag.endswith("depend")}
    assert {"roscpp", "sensor_msgs", "message_generation", "message_runtime", "catkin"} <= deps
    assert not any("pcl" in d for d in deps)


@pytest.mark.skipif(shutil.which("cmake") is None, reason="cmake not available")
def test_catkin_cmake_builds_both_nodes_against_the_stand_in(tmp_path):
    from cones_perception_b200.build import GPU_LIB, build_gpu
    build_gpu()
    ref = _fake_reference(tmp_path)
    bdir = tmp_path / "build"
    prefix = tmp_path / "install"
    subprocess.run(["cmake", "-S", SHELL, "-B", str(bdir), f"-Dcatkin_DIR={STUB}", f"-DCONES_REFERENCE_DIR={ref}",
                    f"-DCONESGPU_PREBUILT={GPU_LIB}", f"-DCMAKE_INSTALL_PREFIX={prefix}", "-DCMAKE_BUILD_TYPE=Release"],
                   check=True, capture_output=True)
    subprocess.run(["cmake", "--build", str(bdir), "-j", "4"], check=True, capture_output=True)
    for exe in ("cone_detection", "ground_removal"):              # the reference's executable names
        assert (bdir / exe).exists(), exe
    calls = (bdir / "catkin_stub_calls.txt").read_text()
    assert "add_service_files:ClassifyColorSrv.srv" in calls and "generate_messages" in calls
    assert "catkin_package" in calls and "conesgpu" in calls and "message_runtime" in calls
    subprocess.run(["cmake", "--install", str(bdir)], check=True, capture_output=True)
    share = prefix / "share" / "cones_perception"
    assert (prefix / "lib" / "cones_perception" / "cone_detection").exists()
    assert (prefix / "lib" / "cones_perception" / "ground_removal").exists()
    assert (prefix / "lib" / "cones_perception" / "color_classifier_server.py").exists()   # the reference's own service
    assert (share / "launch" / "cones_detection.launch").exists()                         # reference launch, unchanged
    assert (share / "launch" / "cones_detection_gpu.launch").exists()                     # the one added launch file
    assert (share / "config" / "ground_removal_params.yaml").exists()
    # without a reference checkout the configure step must stop with a clear message
    r = subprocess.run(["cmake", "-S", SHELL, "-B", str(tmp_path / "b2"), f"-Dcatkin_DIR={STUB}",
                        f"-DCONESGPU_PREBUILT={GPU_LIB}"], capture_output=True, text=True)
    assert r.returncode != 0 and "CONES_REFERENCE_DIR" in r.stderr


@pytest.mark.skipif(shutil.which("cmake") is None or shutil.which("nvcc") is None, reason="cmake / nvcc not available")
def test_catkin_cmake_configures_the_cuda_library_for_sm_100a(tmp_path):
    ref = _fake_reference(tmp_path)
    bdir = tmp_path / "build"
    subprocess.run(["cmake", "-S", SHELL, "-B", str(bdir), f"-Dcatkin_DIR={STUB}", f"-DCONES_REFERENCE_DIR={ref}"],
                   check=True, capture_output=True)
    # the generated build rules compile pipeline.cu with the sm_100a gencode flag
    hits = subprocess.run(["grep", "-rl", "arch=compute_100a,code=sm_100a", str(bdir)], capture_output=True, text=True)
    assert hits.stdout.strip(), "no build rule carries -gencode arch=compute_100a,code=sm_100a"
    rules = subprocess.run(["grep", "-rl", "pipeline.cu", str(bdir)], capture_output=True, text=True)
    assert rules.stdout.strip()
