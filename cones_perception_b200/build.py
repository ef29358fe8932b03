"""Build recipe: compiles libconesgpu.so (sm_100a) and libconesscan.so in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DIR = os.path.join(HERE, "_lib")
GPU_LIB = os.path.join(LIB_DIR, "libconesgpu.so")
SCAN_LIB = os.path.join(HERE, "scangen", "libconesscan.so")
HOST_LIB = os.path.join(LIB_DIR, "libconeshost.so")
ROS_SHELL_LIB = os.path.join(LIB_DIR, "libconesrosshell.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _fresh(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libconesgpu.so cannot be built (there is no CPU fallback)")


def build_gpu(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))]
    srcs.append(os.path.join(ROOT, "include", "conesgpu.h"))
    if not force and _fresh(GPU_LIB, srcs):
        return GPU_LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", GPU_LIB, os.path.join(csrc, "pipeline.cu")]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.run(cmd, check=True)
    return GPU_LIB


def build_scangen(force: bool = False) -> str:
    src = os.path.join(HERE, "scangen", "scan_gen.cpp")
    if not force and _fresh(SCAN_LIB, [src]):
        return SCAN_LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", SCAN_LIB, src], check=True)
    return SCAN_LIB


def build_host(force: bool = False) -> str:
    """C++ host mirror of the reference node classes (host/nodes.hpp) behind a C test surface."""
    hdir = os.path.join(HERE, "host")
    srcs = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir))] + [os.path.join(ROOT, "include", "conesgpu.h")]
    if not force and _fresh(HOST_LIB, srcs + [GPU_LIB]):
        return HOST_LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", HOST_LIB,
                    os.path.join(hdir, "host_c_api.cpp"), "-L" + LIB_DIR, "-lconesgpu", "-Wl,-rpath,$ORIGIN"],
                   check=True)
    return HOST_LIB


def build_ros_shell(force: bool = False) -> str:
    """Test infrastructure: the drop-in node shells (ros_shell/*.cpp) compiled against the stand-in ROS surface and
    message pump of oracle/ref_shim, so they can be run next to the reference's own nodes (ROS is not installed)."""
    shell = os.path.join(ROOT, "ros_shell")
    shim = os.path.join(ROOT, "oracle", "ref_shim")
    srcs = [os.path.join(shell, f) for f in sorted(os.listdir(shell))]
    srcs += [os.path.join(shim, "pump_impl.hpp"), os.path.join(shim, "include", "shim_core.hpp"),
             os.path.join(HERE, "host", "nodes.hpp"), os.path.join(HERE, "host", "pointcloud2.hpp"),
             os.path.join(ROOT, "include", "conesgpu.h")]
    if not force and _fresh(ROS_SHELL_LIB, srcs + [GPU_LIB]):
        return ROS_SHELL_LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", shim, "-I", os.path.join(shim, "include"),
                    "-I", os.path.join(ROOT, "include"), "-I", shell, "-o", ROS_SHELL_LIB,
                    os.path.join(shell, "shim_harness.cpp"), "-L" + LIB_DIR, "-lconesgpu", "-Wl,-rpath,$ORIGIN"],
                   check=True)
    return ROS_SHELL_LIB


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_gpu(force, verbose)
    build_scangen(force)
    build_host(force)
    build_ros_shell(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(GPU_LIB)
    print(SCAN_LIB)
    print(HOST_LIB)
    print(ROS_SHELL_LIB)
