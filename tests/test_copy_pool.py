"""csrc/copy_pool.hpp (parallel, non-temporal staging copy of pageable clouds) on the CPU: compiled with g++ into
a small harness and run with 0, 1 and 3 helper threads."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("workers", [0, 1, 3])
def test_copy_pool_moves_every_byte(tmp_path, workers):
    exe = tmp_path / "copy_pool_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(exe), os.path.join(HERE, "cpp", "copy_pool_test.cpp")],
                   check=True)
    out = subprocess.run([str(exe), str(workers)], check=True, capture_output=True, text=True, timeout=300).stdout
    assert out.strip() == f"ok workers={workers}"
