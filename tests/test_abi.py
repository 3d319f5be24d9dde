"""CPU tier: the C-ABI library loads, exports every symbol include/rfwb200.h declares, keeps the wire
layouts of the reference's #[repr(C)] structs (the contract of backends/metal/src/lib.rs:270-348), and
fails loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from rfw_rs_b200 import backend, wire

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(backend.LIB_PATH):
        backend.build_library()
    return backend.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "rfwb200.h")).read()
    names = re.findall(r"RFWB200_API[^;]*?\b(rfwb200_\w+)\s*\(", hdr)
    assert len(names) >= 40
    raw = C.CDLL(backend.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.rfwb200_version()


def test_wire_layout_matches_c_header(tmp_path):
    # compile a tiny C program that prints sizeof/offsetof from the header and compare with the numpy dtypes
    fields = {
        "RfwRTTriangle": ["vertex0", "u0", "vertex1", "vertex2", "normal", "v0", "n0", "n1", "n2", "id", "tangent0", "tangent1", "tangent2", "light_id", "mat_id", "lod", "area"],
        "RfwDeviceMaterial": ["color", "absorption", "specular", "parameters", "flags", "diffuse_map", "emissive_map", "sheen_map"],
        "RfwCameraView3D": ["pos", "right", "up", "p1", "direction", "lens_size", "spread_angle", "inv_width", "inv_height", "fov", "custom0"],
        "RfwAreaLight": ["position", "energy", "normal", "area", "vertex0", "inst_idx", "vertex1", "mesh_id", "radiance", "vertex2"],
        "RfwSpotLight": ["position", "cos_inner", "radiance", "cos_outer", "direction", "energy"],
        "RfwVertex3D": ["vertex", "normal", "mat_id", "uv", "tangent"],
        "RfwRay": ["origin", "tmin", "direction", "tmax"],
        "RfwHit": ["inst", "prim", "t", "u", "v"],
    }
    src = ['#include <stdio.h>', '#include "rfwb200.h"', "int main(void){"]
    for st in wire.EXPECTED_SIZES:
        src.append(f'printf("{st} %zu\\n", sizeof({st}));')
    for st, fs in fields.items():
        for f in fs:
            src.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    # the ctypes repacks of the non-POD reference types and of the stats / config records
    repacks = {"RfwAabb": wire.CAabb, "RfwMeshData3D": wire.CMeshData3D, "RfwInstancesData3D": wire.CInstancesData3D, "RfwTextureData": wire.CTextureData,
               "RfwB200Config": wire.CConfig, "RfwBuildStats": wire.CBuildStats, "RfwTraceStats": wire.CTraceStats, "RfwRenderStats": wire.CRenderStats,
               "RfwSkinData": wire.CSkinData}
    for st, ct in repacks.items():
        src.append(f'printf("sizeof:{st} %zu\\n", sizeof({st}));')
        for name, *_ in ct._fields_:
            src.append(f'printf("{st}.{name} %zu\\n", offsetof({st}, {name}));')
    src.append("return 0;}")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for st, (dt, size) in wire.EXPECTED_SIZES.items():
        assert int(out[st]) == size == dt.itemsize, st
    for st, fs in fields.items():
        dt = wire.EXPECTED_SIZES[st][0]
        for f in fs:
            assert int(out[f"{st}.{f}"]) == dt.fields[f][1], (st, f)
    # ctypes repacks agree with the header too: size and every field offset
    assert C.sizeof(wire.CAabb) == 32
    for st, ct in repacks.items():
        assert int(out[f"sizeof:{st}"]) == C.sizeof(ct), st
        for name, *_ in ct._fields_:
            assert int(out[f"{st}.{name}"]) == getattr(ct, name).offset, (st, name)


def test_create_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(backend.RfwError) as ei:
        backend.B200Backend(64, 64)
    assert "no CUDA device" in str(ei.value)


def test_null_handle_is_an_error_not_a_crash(lib):
    assert lib.rfwb200_synchronize(None) != 0
    assert b"null backend handle" in lib.rfwb200_last_error()
    assert lib.rfwb200_trace_closest(None, None, 0, None) != 0
