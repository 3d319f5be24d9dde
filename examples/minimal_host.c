/* examples/minimal_host.c — the backend driven from plain C through include/rfwb200.h, the way a host that is neither Rust nor
 * Python would (C99, no CUDA headers, no C++): one quad of two triangles, an identity instance, a material, synchronize, a closest-hit
 * ray, the rest of TIntersector (intersect_t / depth_test), and a 1 spp render.
 *
 *   gcc -std=c99 -Iinclude examples/minimal_host.c -o minimal_host -Lrfw_rs_b200 -lrfwb200 -Wl,-rpath,$PWD/rfw_rs_b200 -lm
 *
 * Exit code 0 = every answer is what geometry says it is.  tests/test_abi.py compiles it on the CPU tier; the GPU tier runs it. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "rfwb200.h"

static void tri(RfwRTTriangle* t, const float* a, const float* b, const float* c, int id) {
    memset(t, 0, sizeof(*t));
    memcpy(t->vertex0, a, 12); memcpy(t->vertex1, b, 12); memcpy(t->vertex2, c, 12);
    t->normal[2] = -1.0f;                                   /* facing the camera at -z */
    t->n0[2] = t->n1[2] = t->n2[2] = -1.0f;
    t->id = id; t->light_id = -1; t->mat_id = 0; t->area = 2.0f;
}

int main(void) {
    RfwB200Config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.width = 64; cfg.height = 48; cfg.max_depth = 3; cfg.clamp_value = 10.0f; cfg.world = 1;
    cfg.sky[0] = 0.2f; cfg.sky[1] = 0.3f; cfg.sky[2] = 0.5f;
    void* h = NULL;
    if (rfwb200_create(&cfg, &h) != RFWB200_OK) { fprintf(stderr, "create: %s\n", rfwb200_last_error()); return 2; }

    const float p00[3] = {-1, -1, 2}, p10[3] = {1, -1, 2}, p11[3] = {1, 1, 2}, p01[3] = {-1, 1, 2};
    RfwRTTriangle tris[2];
    tri(&tris[0], p00, p10, p11, 0);
    tri(&tris[1], p00, p11, p01, 1);
    RfwMeshData3D mesh;
    memset(&mesh, 0, sizeof(mesh));
    mesh.triangles = tris; mesh.num_triangles = 2;
    mesh.bounds.min[0] = -1; mesh.bounds.min[1] = -1; mesh.bounds.min[2] = 2; mesh.bounds.max[0] = 1; mesh.bounds.max[1] = 1; mesh.bounds.max[2] = 2;
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    RfwInstancesData3D inst;
    memset(&inst, 0, sizeof(inst));
    inst.matrices = identity; inst.num_instances = 1; inst.local_aabb = mesh.bounds;
    RfwDeviceMaterial mat;
    memset(&mat, 0, sizeof(mat));
    mat.color[0] = mat.color[1] = mat.color[2] = 0.8f; mat.color[3] = 1.0f;
    mat.diffuse_map = mat.normal_map = mat.metallic_roughness_map = mat.emissive_map = mat.sheen_map = -1;
    int rc = 0;
    rc |= rfwb200_set_3d_mesh(h, 0, &mesh);
    rc |= rfwb200_set_3d_instances(h, 0, &inst);
    rc |= rfwb200_set_materials(h, &mat, 1, NULL);
    rc |= rfwb200_synchronize(h);
    if (rc) { fprintf(stderr, "scene: %s\n", rfwb200_last_error()); return 3; }

    RfwRay rays[2];
    memset(rays, 0, sizeof(rays));
    rays[0].origin[0] = 0.25f; rays[0].origin[1] = -0.5f; rays[0].direction[2] = 1.0f; rays[0].tmin = 1e-4f; rays[0].tmax = 1e26f;   /* hits triangle 0 at t = 2 */
    rays[1] = rays[0]; rays[1].origin[0] = 3.0f;                                                                                  /* misses */
    RfwHit hits[2];
    float t[2], t2[2];
    uint32_t depth[2], occ[2];
    rc |= rfwb200_trace_closest(h, rays, 2, hits);
    rc |= rfwb200_trace_any(h, rays, 2, occ);
    rc |= rfwb200_intersect_t(h, rays, 2, t);
    rc |= rfwb200_depth_test(h, rays, 2, t2, depth);
    if (rc) { fprintf(stderr, "trace: %s\n", rfwb200_last_error()); return 4; }
    printf("ray 0: inst %d prim %d t %.6f (u %.4f v %.4f), occluded %u, intersect_t %.6f, depth %u | ray 1: prim %d t %g occluded %u intersect_t %g\n",
           hits[0].inst, hits[0].prim, hits[0].t, hits[0].u, hits[0].v, occ[0], t[0], depth[0], hits[1].prim, hits[1].t, occ[1], t[1]);
    int bad = 0;
    bad |= !(hits[0].inst == 0 && hits[0].prim == 0 && fabsf(hits[0].t - 2.0f) < 1e-5f && occ[0] == 1 && fabsf(t[0] - 2.0f) < 1e-5f && fabsf(t2[0] - 2.0f) < 1e-5f);
    bad |= !(hits[1].inst == -1 && hits[1].prim == -1 && hits[1].t == rays[1].tmax && occ[1] == 0 && t[1] == -1.0f && t2[1] == rays[1].tmax);

    RfwCameraView3D view;
    memset(&view, 0, sizeof(view));
    view.pos[2] = -1.0f;                                    /* pinhole at (0, 0, -1) looking down +z: image plane z = 0, x, y in [-1, 1] */
    view.right[0] = 2.0f; view.up[1] = 2.0f; view.p1[0] = -1.0f; view.p1[1] = -1.0f; view.p1[2] = 0.0f;
    view.direction[2] = 1.0f; view.inv_width = 1.0f / 64.0f; view.inv_height = 1.0f / 48.0f; view.fov = 90.0f;
    rc = rfwb200_render(h, &view, RFW_RENDER_DEFAULT);
    static float image[64 * 48 * 4];
    rc |= rfwb200_read_output(h, image);
    if (rc) { fprintf(stderr, "render: %s\n", rfwb200_last_error()); return 5; }
    /* no lights: the quad is black (sqrt(0)), the sky around it is sqrt(sky) */
    const float* centre = image + 4 * (24 * 64 + 32);
    const float* corner = image + 4 * (1 * 64 + 1);
    printf("centre pixel %.3f %.3f %.3f, corner pixel %.3f %.3f %.3f, samples %u, version %s\n", centre[0], centre[1], centre[2], corner[0], corner[1], corner[2],
           rfwb200_sample_count(h), rfwb200_version());
    bad |= !(fabsf(corner[0] - sqrtf(0.2f)) < 1e-3f && fabsf(corner[2] - sqrtf(0.5f)) < 1e-3f && rfwb200_sample_count(h) == 1);
    rfwb200_destroy(h);
    return bad ? 1 : 0;
}
