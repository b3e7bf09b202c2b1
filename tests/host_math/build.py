"""Builds tests/host_math/_host_math.so (g++, no CUDA) -- a TEST TOOL that lets the CPU suite check the
kernels' per-element math header against the oracle.  Never loaded by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_host_math.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "host_math.cpp")
    hdr = os.path.join(ROOT, "pipeline_pointcloud_b200", "csrc", "b2s_math.cuh")
    if (not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return SO
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-x", "c++",
           "-I", os.path.dirname(hdr), src, "-o", SO]
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True))
