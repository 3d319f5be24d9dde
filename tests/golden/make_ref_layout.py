#!/usr/bin/env python
"""Generates tests/golden/ref_metal_layout.json: sizeof / offsetof of the C structs the REFERENCE's own FFI boundary declares
(backends/metal/cpp/src/structs.h, library.h) — the C side of the reference's only ABI test, `test_layout`
(backends/metal/src/lib.rs:270-348), which asserts that these C structs have the sizes of the Rust #[repr(C)] types
(RTTriangle, CameraView3D, DeviceMaterial, Vertex3D, Aabb, VertexMesh, JointData).

The reference headers are compiled WHERE THEY LIE under /root/reference (nothing is copied) with a three-typedef stand-in for
Apple's <simd/simd.h>; only the resulting numbers are committed.  Run in the build container:  python tests/golden/make_ref_layout.py
"""
import json
import os
import subprocess
import sys
import tempfile

REF = os.environ.get("RFW_REFERENCE", "/root/reference")
HDR_DIR = os.path.join(REF, "backends", "metal", "cpp", "src")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_metal_layout.json")

FIELDS = {
    "RTTriangle": ["vertex0", "u0", "vertex1", "u1", "vertex2", "u2", "normal", "v0", "n0", "v1", "n1", "v2", "n2", "id", "tangent0", "tangent1", "tangent2",
                   "light_id", "mat_id", "lod", "area"],
    "CameraView3D": ["pos", "right", "up", "p1", "direction", "lens_size", "spread_angle", "epsilon", "inv_width", "inv_height", "near_plane", "far_plane",
                     "aspect_ratio", "fov", "custom0", "custom1"],
    "DeviceMaterial": ["c_r", "a_r", "s_r", "params_x", "flags", "diffuse_map", "normal_map", "metallic_roughness_map", "emissive_map", "sheen_map"],
    "Vertex3D": ["v_x", "n_x", "mat_id", "u", "t_x"],
    "Aabb": ["bmin", "bmax"],
    "VertexRange": ["bounds", "first", "last", "mat_id", "padding"],
    "JointData": ["j_x", "weight"],
}

SIMD_STUB = """#pragma once
typedef float simd_float2 __attribute__((vector_size(8)));
typedef float simd_float4 __attribute__((vector_size(16)));
typedef struct { simd_float4 columns[4]; } simd_float4x4;
"""


def main():
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "simd"))
        open(os.path.join(tmp, "simd", "simd.h"), "w").write(SIMD_STUB)
        lines = ["#include <cstdio>", "#include <cstddef>", '#include "library.h"', "int main() {"]
        for st, fs in FIELDS.items():
            lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
            for f in fs:
                lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
        lines.append("return 0; }")
        src = os.path.join(tmp, "layout.cpp")
        open(src, "w").write("\n".join(lines))
        exe = os.path.join(tmp, "layout")
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-w", "-I", tmp, "-I", HDR_DIR, src, "-o", exe])
        out = {k: int(v) for k, v in (line.split() for line in subprocess.check_output([exe]).decode().splitlines())}
    doc = {"source": "backends/metal/cpp/src/{structs.h,library.h} of the reference, compiled with g++ and a stand-in <simd/simd.h>", "layout": out}
    json.dump(doc, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(out), "entries")


if __name__ == "__main__":
    sys.exit(main())
