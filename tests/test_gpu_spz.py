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


# ---- the native file-level PLY -> SPZ entry point and the converter command line (csrc/spz_file.cu, splat_converter_main.cpp) ----
CONVERTER = os.path.join(os.path.dirname(HERE), "pipeline_pointcloud_b200", "splat_converter")


def test_native_ply_to_spz_matches_reference_golden_file(tmp_path):
    out = str(tmp_path / "x.spz")
    assert export.convert_ply_to_spz_native(os.path.join(HERE, "golden", "splat_small.ply"), out) == out
    assert open(out, "rb").read() == open(os.path.join(HERE, "golden", "splat_small.spz"), "rb").read()


@pytest.mark.parametrize("rows,n", [(0, 1000), (3, 70001), (8, 4097), (15, 120000)])
def test_native_ply_to_spz_equals_python_path(tmp_path, rows, n):
    sp = random_splats(n, rows, seed=40 + rows)
    ply = str(tmp_path / "splat.ply")
    export.write_ply(ply, export.ply_channels(sp))
    export.convert_ply_to_spz(ply)                                # Python-driven: read_ply + b2s_spz_pack + zlib
    via_python = open(str(tmp_path / "splat.spz"), "rb").read()
    export.convert_ply_to_spz_native(ply, str(tmp_path / "native.spz"))
    assert open(str(tmp_path / "native.spz"), "rb").read() == via_python


@pytest.mark.skipif(not os.path.exists(CONVERTER), reason="pipeline_pointcloud_b200/splat_converter not built")
def test_converter_command_line_both_directions(tmp_path):
    """`splat_converter splat.ply` (what the Spz-Export component runs, main.py:1532-1545) and `splat_converter splat.spz`:
    same files, same message, same exit codes as the reference binary."""
    sp = random_splats(50000, 15, seed=61)
    ply = str(tmp_path / "splat.ply")
    export.write_ply(ply, export.ply_channels(sp))
    r = subprocess.run([CONVERTER, ply], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == f'Converted {ply} to "{tmp_path / "splat.spz"}"', (r.stdout, r.stderr)
    mine_spz = open(str(tmp_path / "splat.spz"), "rb").read()
    os.rename(ply, str(tmp_path / "orig.ply"))
    r = subprocess.run([CONVERTER, str(tmp_path / "splat.spz")], capture_output=True, text=True)
    assert r.returncode == 0 and os.path.exists(ply)
    mine_ply = open(ply, "rb").read()
    assert subprocess.run([CONVERTER, str(tmp_path / "nothing.ply")], capture_output=True).returncode == 1
    assert subprocess.run([CONVERTER, str(tmp_path / "orig.ply"), "extra"], capture_output=True).returncode == 1
    other = tmp_path / "scene.obj"
    other.write_text("x")
    assert subprocess.run([CONVERTER, str(other)], capture_output=True).returncode == 1
    if os.path.exists(REF_BIN):                                   # the reference binary on the same inputs
        ref_dir = tmp_path / "ref"
        ref_dir.mkdir()
        os.link(str(tmp_path / "orig.ply"), str(ref_dir / "splat.ply"))
        r = subprocess.run([REF_BIN, str(ref_dir / "splat.ply")], capture_output=True, text=True)
        assert r.returncode == 0 and open(str(ref_dir / "splat.spz"), "rb").read() == mine_spz
        os.remove(str(ref_dir / "splat.ply"))
        subprocess.run([REF_BIN, str(ref_dir / "splat.spz")], check=True, capture_output=True)
        assert open(str(ref_dir / "splat.ply"), "rb").read() == mine_ply
