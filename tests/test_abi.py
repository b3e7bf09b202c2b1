"""CPU: the C-ABI library loads and exports every symbol include/b2s.h declares; the host wrapper fails loudly
without a GPU (no compute calls here)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2s.h")).read()
    return sorted(set(re.findall(r"B2S_API\s+[\w\s\*]+?\b(b2s_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from pipeline_pointcloud_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 16
    handle = _lib.lib()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/b2s.h but not exported by libb2s.so"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype"
    assert set(_lib.SYMBOLS) == set(names)
    assert handle.b2s_version() == 100
    assert handle.b2s_last_error() is not None


def test_argument_errors_do_not_need_a_gpu():
    import ctypes
    from pipeline_pointcloud_b200 import _lib
    h = _lib.lib()
    assert h.b2s_project_fwd(None, None) == -1
    assert b"null args" in h.b2s_last_error()
    a = _lib.CompositeFwdArgs(C=1, width=32, height=32, tile_size=8, tile_w=4, tile_h=4, channels=3)
    assert h.b2s_composite_fwd(ctypes.byref(a), None) == -1
    assert h.b2s_isect_scan_workspace_bytes(1_000_000) > 0


def test_no_cpu_fallback():
    import pipeline_pointcloud_b200 as p
    from pipeline_pointcloud_b200._lib import B2SError
    with pytest.raises(B2SError):
        p.rasterization(torch.zeros(4, 3), torch.zeros(4, 4), torch.ones(4, 3), torch.ones(4), torch.zeros(4, 3),
                        torch.eye(4)[None], torch.eye(3)[None], 32, 32)
    with pytest.raises(B2SError):
        p.isect_tiles(torch.zeros(1, 4, 2), torch.zeros(1, 4, dtype=torch.int32), torch.zeros(1, 4), 16, 2, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pipeline_pointcloud_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_reference_codec_entry_point_is_exported(tmp_path):
    """``ConvertSpzToPly`` -- the C symbol the reference's codec library exports (spz/src/spz_converter_api.cpp:7) -- is
    exported under the same name and signature; its file and header checks run on the host (no GPU needed to fail)."""
    import ctypes
    from pipeline_pointcloud_b200 import _lib
    h = _lib.lib()
    fn = h.ConvertSpzToPly
    fn.restype, fn.argtypes = ctypes.c_bool, [ctypes.c_char_p, ctypes.c_char_p]
    out = str(tmp_path / "o.ply").encode()
    assert fn(str(tmp_path / "missing.spz").encode(), out) is False
    assert b"does not exist" in h.b2s_last_error()
    bad = tmp_path / "bad.spz"
    bad.write_bytes(b"this is not a gzip container")
    assert h.b2s_convert_spz_to_ply(str(bad).encode(), out) == -1 and b"gzip" in h.b2s_last_error()
    import gzip
    import struct
    wrong = tmp_path / "wrong.spz"                      # a gzip container whose payload is not an SPZ header
    wrong.write_bytes(gzip.compress(struct.pack("<IIIBBBB", 0x12345678, 2, 1, 0, 12, 0, 0) + bytes(64)))
    assert h.b2s_convert_spz_to_ply(str(wrong).encode(), out) == -1 and b"magic" in h.b2s_last_error()
    short = tmp_path / "short.spz"                      # a valid header that promises more points than the payload holds
    short.write_bytes(gzip.compress(struct.pack("<IIIBBBB", 0x5053474E, 2, 1000, 3, 12, 0, 0) + bytes(64)))
    assert h.b2s_convert_spz_to_ply(str(short).encode(), out) == -1 and b"shorter" in h.b2s_last_error()
    assert not os.path.exists(out.decode())
    if not torch.cuda.is_available():                   # a well-formed file without a device: a CUDA error code, no crash
        golden = os.path.join(ROOT, "tests", "golden", "splat_small.spz").encode()
        assert h.b2s_convert_spz_to_ply(golden, out) > 0


def test_native_ply_reader_follows_the_reference_header_rules(tmp_path):
    """b2s_convert_ply_to_spz parses the header as loadSplatFromPly does (load-spz.cc:581-645): every rejection happens on the
    host, before any CUDA call."""
    from pipeline_pointcloud_b200 import _lib
    h = _lib.lib()
    out = str(tmp_path / "o.spz").encode()
    fields = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
              "rot_0", "rot_1", "rot_2", "rot_3"]

    def ply(first="ply", fmt="format binary_little_endian 1.0", n=2, props=None, body=None, extra=()):
        props = fields if props is None else props
        head = [first, fmt, *extra, f"element vertex {n}", *[f"property float {k}" for k in props], "end_header"]
        data = bytes(4 * max(n, 0) * len(props)) if body is None else body
        path = tmp_path / "t.ply"
        path.write_bytes(("\n".join(head) + "\n").encode() + data)
        return str(path).encode()

    def rejected(path, needle):
        assert h.b2s_convert_ply_to_spz(path, out) == -1, needle
        assert needle in h.b2s_last_error(), h.b2s_last_error()

    rejected(str(tmp_path / "missing.ply").encode(), b"unable to open")
    rejected(ply(first="plx"), b"not a .ply file")
    rejected(ply(fmt="format ascii 1.0"), b"unsupported .ply format")
    rejected(ply(n=0), b"invalid vertex count")
    rejected(ply(n=10 * 1024 * 1024 + 1, body=b""), b"invalid vertex count")
    rejected(ply(props=[k for k in fields if k != "rot_3"]), b"missing field: rot_3")
    rejected(ply(props=fields + ["f_rest_0", "f_rest_1"]), b"f_rest_*")
    rejected(ply(body=bytes(8)), b"shorter")
    bad = tmp_path / "u.ply"
    bad.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty uchar red\nend_header\n\x00")
    rejected(str(bad).encode(), b"unsupported property data type")
    assert not os.path.exists(out.decode())
    if not torch.cuda.is_available():       # a well-formed file (with a comment line, as the reference tolerates): CUDA error, no crash
        assert h.b2s_convert_ply_to_spz(ply(extra=("comment made by a test",)), out) > 0
