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


def test_wire_layout_matches_the_reference_ffi_headers(tmp_path):
    """Golden fixture from the reference itself: sizeof / offsetof of the C structs of its own FFI boundary
    (backends/metal/cpp/src/structs.h + library.h, compiled where they lie by tests/golden/make_ref_layout.py) — the C side of
    the reference's only ABI test (`test_layout`, backends/metal/src/lib.rs:270-348), which pins them to the Rust #[repr(C)]
    types.  include/rfwb200.h must lay the same seven structs out identically, field by field."""
    import json

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_metal_layout.json")))["layout"]
    # reference C struct -> (our struct, {reference field: our field})
    same = lambda names: {n: n for n in names}
    mapping = {
        "RTTriangle": ("RfwRTTriangle", same(["vertex0", "u0", "vertex1", "u1", "vertex2", "u2", "normal", "v0", "n0", "v1", "n1", "v2", "n2", "id", "tangent0", "tangent1",
                                             "tangent2", "light_id", "mat_id", "lod", "area"])),
        "CameraView3D": ("RfwCameraView3D", same(["pos", "right", "up", "p1", "direction", "lens_size", "spread_angle", "epsilon", "inv_width", "inv_height", "near_plane",
                                                 "far_plane", "aspect_ratio", "fov", "custom0", "custom1"])),
        "DeviceMaterial": ("RfwDeviceMaterial", {"c_r": "color", "a_r": "absorption", "s_r": "specular", "params_x": "parameters", "flags": "flags", "diffuse_map": "diffuse_map",
                                                 "normal_map": "normal_map", "metallic_roughness_map": "metallic_roughness_map", "emissive_map": "emissive_map",
                                                 "sheen_map": "sheen_map"}),
        "Vertex3D": ("RfwVertex3D", {"v_x": "vertex", "n_x": "normal", "mat_id": "mat_id", "u": "uv", "t_x": "tangent"}),
        "Aabb": ("RfwAabb", {"bmin": "min", "bmax": "max"}),
        "VertexRange": ("RfwVertexMesh", same(["bounds", "first", "last", "mat_id", "padding"])),
        "JointData": ("RfwJointData", {"j_x": "joint", "weight": "weight"}),
    }
    src = ['#include <stdio.h>', '#include "rfwb200.h"', "int main(void){"]
    for ref, (ours, fields) in mapping.items():
        src.append(f'printf("{ref} %zu\\n", sizeof({ours}));')
        for rf, of in fields.items():
            src.append(f'printf("{ref}.{rf} %zu\\n", offsetof({ours}, {of}));')
    src.append("return 0;}")
    c = tmp_path / "ref_layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "ref_layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = {k: int(v) for k, v in (line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())}
    assert len(out) >= 60
    for k, v in out.items():
        assert gold[k] == v, (k, gold[k], v)
    # the fixture covers every field the reference header declares for these structs (the generator lists them all)
    assert set(out) == set(gold)


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


def _build_c_host(tmp_path):
    exe = str(tmp_path / "minimal_host")
    lib_dir = os.path.dirname(backend.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "minimal_host.c"), "-o", exe,
                           "-L" + lib_dir, "-lrfwb200", "-Wl,-rpath," + lib_dir, "-lm"])
    return exe


def test_plain_c_host_compiles_and_links(lib, tmp_path):
    """examples/minimal_host.c — a host in plain C99 (no C++, no CUDA headers): include/rfwb200.h is a C header and librfwb200.so links
    from C.  Without a GPU the program must fail loudly at rfwb200_create (exit code 2), not crash."""
    import torch

    exe = _build_c_host(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 2 and "create:" in r.stderr, (r.returncode, r.stderr[-300:])


@pytest.mark.gpu
def test_plain_c_host_runs(lib, tmp_path):
    """The same program on the GPU box: quad, instance, material, synchronize, closest / any-hit, intersect_t, depth_test and a 1 spp render, all from C."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([_build_c_host(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-600:], r.stderr[-600:])
    assert "prim 0 t 2.000000" in r.stdout
