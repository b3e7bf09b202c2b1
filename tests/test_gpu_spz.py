"""GPU: the CUDA SPZ packer (csrc/spz.cu, through export.py) byte for byte against the golden file written by the
reference's own converter, the numpy oracle, and the compiled reference itself (oracle/_ref/splat_converter)."""
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import spz_oracle as S
from pipeline_pointcloud_b200 import export
from tests.test_spz_cpu import REF_BIN, random_splats

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_cuda_packer_matches_reference_golden_file(tmp_path):
    ply = os.path.join(HERE, "golden", "splat_small.ply")
    gold = open(os.path.join(HERE, "golden", "splat_small.spz"), "rb").read()
    sp = export.splats_from_ply(ply)
    assert export.spz_payload(sp) == S.gunzip_bytes(gold)
    out = str(tmp_path / "x.spz")
    export.write_spz(out, sp)
    assert open(out, "rb").read() == gold                    # the whole file, gzip container included


@pytest.mark.parametrize("rows,n", [(0, 1000), (3, 70001), (8, 4097), (15, 300000)])
def test_cuda_packer_matches_oracle(rows, n):
    sp = random_splats(n, rows, seed=10 + rows)
    sp["means"][:50] *= 100.0                                 # beyond the 24-bit range: wraps exactly as the reference
    sp["scales"][50:100] = torch.linspace(-12, 8, 150).reshape(50, 3)
    sp["opacities"][100:150] = torch.linspace(-30, 30, 50)
    pos, sc, rot, al, col, sh = (sp["means"].numpy(), sp["scales"].numpy(), sp["quats"][:, [1, 2, 3, 0]].numpy(),
                                 sp["opacities"].numpy(), sp["sh0"][:, 0].numpy(), sp["shN"].numpy())
    ref = S.serialize(S.pack_gaussians(pos, sc, rot, al, col, sh), n, S.degree_for_dim(rows))
    got = export.spz_payload({k: v.cuda() for k, v in sp.items()})
    a, b = np.frombuffer(got, np.uint8), np.frombuffer(ref, np.uint8)
    assert a.shape == b.shape and int((a != b).sum()) == 0


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference (oracle/_ref/splat_converter) not built")
def test_cuda_converter_matches_compiled_reference(tmp_path):
    sp = random_splats(200000, 15, seed=99)
    ply = str(tmp_path / "splat.ply")
    export.write_ply(ply, export.ply_channels(sp))
    assert export.main([ply]) == 0                            # `splat_converter splat.ply` -> splat.spz, on the GPU
    mine = open(str(tmp_path / "splat.spz"), "rb").read()
    os.rename(str(tmp_path / "splat.spz"), str(tmp_path / "mine.spz"))
    subprocess.run([REF_BIN, ply], check=True, stdout=subprocess.DEVNULL)
    ref = open(str(tmp_path / "splat.spz"), "rb").read()
    assert mine == ref


# ---- the decode direction (csrc/spz.cu spz_unpack_kernel, csrc/spz_file.cu ConvertSpzToPly) ----
from tests.test_spz_cpu import oracle_ply_from_spz


def test_cuda_decoder_matches_reference_golden_ply(tmp_path):
    """The native file-level entry point on the golden pair the reference binary wrote (spz -> ply), whole file."""
    out = str(tmp_path / "x.ply")
    assert export.convert_spz_to_ply(os.path.join(HERE, "golden", "splat_small.spz"), out) == out
    assert open(out, "rb").read() == open(os.path.join(HERE, "golden", "splat_small_decoded.ply"), "rb").read()


@pytest.mark.parametrize("rows,n", [(0, 1000), (3, 70001), (8, 4097), (15, 300000)])
def test_cuda_decoder_matches_oracle(rows, n):
    sp = random_splats(n, rows, seed=20 + rows)
    sp["opacities"][:256] = torch.linspace(-12, 12, 256)
    sp["means"][:50] *= 100.0
    payload = export.spz_payload({k: v.cuda() for k, v in sp.items()})
    table = export.ply_table_from_spz_payload(payload).cpu().numpy()
    ref = oracle_ply_from_spz(S.gzip_bytes(payload))
    body = np.frombuffer(ref[ref.index(b"end_header\n") + 11:], dtype="<u4").reshape(n, -1)
    assert table.shape == body.shape and int((table.view(np.uint32) != body).sum()) == 0     # bit patterns: +-inf logits too


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference (oracle/_ref/splat_converter) not built")
def test_cuda_spz_to_ply_matches_compiled_reference(tmp_path):
    import ctypes
    from pipeline_pointcloud_b200 import _lib
    sp = random_splats(200000, 15, seed=77)
    spz = str(tmp_path / "splat.spz")
    export.write_spz(spz, {k: v.cuda() for k, v in sp.items()})
    assert export.main([spz]) == 0                                # `splat_converter splat.spz` -> splat.ply, on the GPU
    mine = open(str(tmp_path / "splat.ply"), "rb").read()
    subprocess.run([REF_BIN, spz], check=True, stdout=subprocess.DEVNULL)
    assert mine == open(str(tmp_path / "splat.ply"), "rb").read()
    fn = _lib.lib().ConvertSpzToPly                               # the reference's exported name (spz_converter_api.cpp:7)
    fn.restype, fn.argtypes = ctypes.c_bool, [ctypes.c_char_p, ctypes.c_char_p]
    assert fn(spz.encode(), str(tmp_path / "again.ply").encode()) is True
    assert open(str(tmp_path / "again.ply"), "rb").read() == mine
    assert fn(str(tmp_path / "missing.spz").encode(), str(tmp_path / "no.ply").encode()) is False
