"""Generates tests/golden/splat_small.ply, tests/golden/splat_small.spz and tests/golden/splat_small_decoded.ply: the .spz is written by the REFERENCE's own
converter (oracle/_ref/splat_converter, compiled by oracle/Makefile from
/root/reference/source/container/src/pipeline/post_processing/spz/src) from the committed .ply, and the decoded .ply by the same binary from that .spz.
Run in the build container:  make -C oracle && python tests/golden/make_spz_golden.py"""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pipeline_pointcloud_b200 import export  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def small_splats(n=400, seed=0):
    g = torch.Generator().manual_seed(seed)
    s = {"means": torch.randn(n, 3, generator=g) * 4.0, "scales": torch.randn(n, 3, generator=g) * 2.0 - 4.0,
         "quats": torch.randn(n, 4, generator=g), "opacities": torch.randn(n, generator=g) * 3.0,
         "sh0": torch.randn(n, 1, 3, generator=g) * 1.5, "shN": torch.randn(n, 15, 3, generator=g) * 0.4}
    # edge cases: clamps on every field, exact .5 ties, negative w, huge coordinates
    s["means"][0] = torch.tensor([2047.99, -2048.0, 0.0001220703125])
    s["means"][1] = torch.tensor([1.5 / 4096, -1.5 / 4096, 2.5 / 4096])
    s["scales"][2] = torch.tensor([-11.0, 6.5, -10.03125])
    s["opacities"][3], s["opacities"][4] = 40.0, -40.0
    s["sh0"][5] = torch.tensor([[5.0, -5.0, 0.0]])
    s["shN"][6, :3] = torch.tensor([[1.5, -1.5, 0.0], [0.99, -1.0, 1.0], [3.5 / 128, 4.5 / 128, -3.5 / 128]])
    s["quats"][7] = torch.tensor([-0.5, 0.5, -0.5, 0.5])
    s["quats"][8] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    return s


if __name__ == "__main__":
    ply = os.path.join(HERE, "splat_small.ply")
    export.write_ply(ply, export.ply_channels(small_splats()))
    conv = os.path.join(ROOT, "oracle", "_ref", "splat_converter")
    subprocess.run([conv, ply], check=True)
    print(os.path.getsize(ply), os.path.getsize(os.path.join(HERE, "splat_small.spz")))
    # the decode direction: the reference binary turns its own .spz back into a .ply (spz/src/main.cpp:29-35)
    import shutil
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(HERE, "splat_small.spz"), os.path.join(d, "a.spz"))
        subprocess.run([conv, os.path.join(d, "a.spz")], check=True)
        shutil.copy(os.path.join(d, "a.ply"), os.path.join(HERE, "splat_small_decoded.ply"))
