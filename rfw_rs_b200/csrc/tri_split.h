// tri_split.h — spatial splits for the BLAS build: triangle PRE-SPLITTING ahead of the Morton sort.
//
// The reference advertises a "Spatial BVH" (backends/gpu-rt/README.md:10; rtbvh's spatial-split builder).  A top-down SBVH does
// not fit a sort-based device build; what does is the pre-splitting of Karras & Aila, "Fast Parallel Construction of
// High-Quality Bounding Volume Hierarchies" (HPG 2013, section 4): a triangle whose box is much larger than the triangle
// warrants — long, thin, diagonal, or simply huge next to its neighbours — is replaced by several REFERENCES to it, each with the
// (tight, clipped) box of the part of the triangle inside one cell of the implicit Morton grid.  The references then go
// through the unchanged build (Morton -> sort -> Karras -> SAH refinement -> wide collapse); the leaf-ordered traversal triangles
// simply repeat a split triangle once per reference.  A ray still meets the whole triangle (the test is not clipped), so hits,
// ids and t are exactly what they are without splits — only fewer boxes overlap.
//
// Bodies are RFW_HD: nvcc compiles them into the kernels of backend.cu, tests/hostemu compiles them for the CPU tier.
#pragma once
#include "hd.h"

namespace rfw {

static constexpr int SPLIT_GRID_BITS = 10;   // implicit Morton grid over the mesh bounds: 1024 cells per axis
static constexpr int SPLIT_MAX_EXTRA = 31;   // at most this many extra references per triangle

struct SplitGrid {
    float3 lo;       // mesh bounds
    float3 scale;    // cells per unit length (0 on a flat axis)
    float3 cell;     // unit length per cell
};

RFW_HD SplitGrid make_split_grid(const float3 lo, const float3 hi) {
    SplitGrid g;
    g.lo = lo;
    const float cells = (float)(1 << SPLIT_GRID_BITS);
    const float3 e = hi - lo;
    g.scale = f3(e.x > 0.0f ? cells / e.x : 0.0f, e.y > 0.0f ? cells / e.y : 0.0f, e.z > 0.0f ? cells / e.z : 0.0f);
    g.cell = f3(e.x / cells, e.y / cells, e.z / cells);
    return g;
}

RFW_HD int split_quantise(float x, float lo, float scale) {
    const float q = (x - lo) * scale;
    const int top = (1 << SPLIT_GRID_BITS) - 1;
    return q <= 0.0f ? 0 : (q >= (float)top ? top : (int)q);
}
// upper ends: a box that ends exactly ON a cell boundary belongs to the cell below it (it does not cross that plane)
RFW_HD int split_quantise_hi(float x, float lo, float scale, int ql) {
    const float q = (x - lo) * scale;
    const int top = (1 << SPLIT_GRID_BITS) - 1;
    int c = q <= 0.0f ? 0 : (q >= (float)(top + 1) ? top : (int)q);
    if (c > 0 && (float)c == q) c--;
    return c < ql ? ql : c;
}

// The most important spatial-median plane of the Morton grid that the box [lo, hi] crosses: axis (0..2) and plane position in
// cells, or axis = -1 when the box lies inside one cell on every axis.  `level` = 3 * (depth of the plane in its axis'
// binary subdivision) + axis: the order in which the Morton code's bits split space (0 = the x median of the whole mesh).
RFW_HD void split_plane(const SplitGrid& g, const float3 lo, const float3 hi, int& axis, int& plane_cell, int& level) {
    axis = -1; plane_cell = 0; level = 3 * SPLIT_GRID_BITS;
    const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z}, gl[3] = {g.lo.x, g.lo.y, g.lo.z}, gs[3] = {g.scale.x, g.scale.y, g.scale.z};
    for (int a = 0; a < 3; a++) {
        const int ql = split_quantise(l[a], gl[a], gs[a]), qh = split_quantise_hi(h[a], gl[a], gs[a], ql);
        if (ql == qh) continue;
        const int b = bfind32((uint32_t)(ql ^ qh));          // highest differing bit
        const int lv = 3 * (SPLIT_GRID_BITS - 1 - b) + a;
        if (lv < level) { level = lv; axis = a; plane_cell = (qh >> b) << b; }
    }
}

RFW_HD float split_box_area(const float3 lo, const float3 hi) {
    const float3 e = hi - lo;
    return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x);
}

// Karras & Aila eq. (4): priority = (2^-level * (A_box - A_ideal))^(1/3); A_ideal = the box area a triangle of this orientation
// cannot go below (the sum of its projected areas, doubled = the components of the edge cross product).  0 when nothing can be gained.
RFW_HD float split_priority(const SplitGrid& g, const float3 v0, const float3 v1, const float3 v2) {
    const float3 lo = f3(fminf(v0.x, fminf(v1.x, v2.x)), fminf(v0.y, fminf(v1.y, v2.y)), fminf(v0.z, fminf(v1.z, v2.z)));
    const float3 hi = f3(fmaxf(v0.x, fmaxf(v1.x, v2.x)), fmaxf(v0.y, fmaxf(v1.y, v2.y)), fmaxf(v0.z, fmaxf(v1.z, v2.z)));
    int axis, cell, level;
    split_plane(g, lo, hi, axis, cell, level);
    if (axis < 0) return 0.0f;
    const float3 e1 = v1 - v0, e2 = v2 - v0;
    const float3 c = f3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
    const float ideal = fabsf(c.x) + fabsf(c.y) + fabsf(c.z);
    const float gain = split_box_area(lo, hi) - ideal;
    if (!(gain > 0.0f)) return 0.0f;
    // areas relative to the mesh: the priority must not depend on the unit of length
    const float3 ge = f3(g.cell.x, g.cell.y, g.cell.z) * (float)(1 << SPLIT_GRID_BITS);
    const float mesh_area = fmaxf(2.0f * (ge.x * ge.y + ge.y * ge.z + ge.z * ge.x), 1e-30f);
    return cbrtf(exp2f(-(float)level) * (gain / mesh_area));
}

RFW_HD void split_grow(float3& lo, float3& hi, const float3 p) {
    lo = f3(fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z));
    hi = f3(fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z));
}
RFW_HD float split_comp(const float3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
RFW_HD void split_set_comp(float3& v, int a, float x) { if (a == 0) v.x = x; else if (a == 1) v.y = x; else v.z = x; }

// Boxes of the parts of triangle (v0, v1, v2) on either side of the plane x[axis] = pos, each intersected with the current box.
// Conservative: an edge / plane intersection is computed in float32, so every box is padded by `pad` afterwards (by the caller).
RFW_HD void split_boxes(const float3 v[3], int axis, float pos, const float3 cur_lo, const float3 cur_hi, float3& l_lo, float3& l_hi, float3& r_lo, float3& r_hi) {
    const float big = 3.0e38f;
    l_lo = r_lo = f3(big, big, big); l_hi = r_hi = f3(-big, -big, -big);
    for (int k = 0; k < 3; k++) {
        const float3 a = v[k], b = v[(k + 1) % 3];
        const float ax = split_comp(a, axis), bx = split_comp(b, axis);
        if (ax <= pos) split_grow(l_lo, l_hi, a);
        if (ax >= pos) split_grow(r_lo, r_hi, a);
        if ((ax < pos && bx > pos) || (ax > pos && bx < pos)) {
            const float t = fminf(fmaxf((pos - ax) / (bx - ax), 0.0f), 1.0f);
            float3 p = f3(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z));
            split_set_comp(p, axis, pos);
            split_grow(l_lo, l_hi, p); split_grow(r_lo, r_hi, p);
        }
    }
    // within the current box, and not across the plane
    l_lo = f3(fmaxf(l_lo.x, cur_lo.x), fmaxf(l_lo.y, cur_lo.y), fmaxf(l_lo.z, cur_lo.z)); l_hi = f3(fminf(l_hi.x, cur_hi.x), fminf(l_hi.y, cur_hi.y), fminf(l_hi.z, cur_hi.z));
    r_lo = f3(fmaxf(r_lo.x, cur_lo.x), fmaxf(r_lo.y, cur_lo.y), fmaxf(r_lo.z, cur_lo.z)); r_hi = f3(fminf(r_hi.x, cur_hi.x), fminf(r_hi.y, cur_hi.y), fminf(r_hi.z, cur_hi.z));
    split_set_comp(l_hi, axis, fminf(split_comp(l_hi, axis), pos));
    split_set_comp(r_lo, axis, fmaxf(split_comp(r_lo, axis), pos));
}

RFW_HD bool split_box_valid(const float3 lo, const float3 hi) { return lo.x <= hi.x && lo.y <= hi.y && lo.z <= hi.z; }

// Replaces one triangle by exactly `count` (>= 1) reference boxes, written to out_lo / out_hi [0, count).  Recursive median
// splits along the Morton grid, the remaining count shared between the halves in proportion to their extent along the split
// axis (Karras & Aila, section 4.2).  Every box is padded by `pad` (absolute) on every side and stays inside the triangle's own box
// (padded alike).  The union of the boxes covers the triangle.
RFW_HD void split_triangle(const SplitGrid& g, const float3 v0, const float3 v1, const float3 v2, int count, float pad, float4* out_lo, float4* out_hi) {
    const float3 v[3] = {v0, v1, v2};
    float3 tlo = f3(fminf(v0.x, fminf(v1.x, v2.x)), fminf(v0.y, fminf(v1.y, v2.y)), fminf(v0.z, fminf(v1.z, v2.z)));
    float3 thi = f3(fmaxf(v0.x, fmaxf(v1.x, v2.x)), fmaxf(v0.y, fmaxf(v1.y, v2.y)), fmaxf(v0.z, fmaxf(v1.z, v2.z)));
    struct Item { float3 lo, hi; int count; };
    Item stack[SPLIT_MAX_EXTRA + 2];
    int sp = 0, emitted = 0;
    if (count > SPLIT_MAX_EXTRA + 1) count = SPLIT_MAX_EXTRA + 1;
    stack[sp++] = Item{tlo, thi, count};
    while (sp > 0) {
        Item it = stack[--sp];
        for (;;) {  // (a part that lies on one side of its plane keeps its count and tries the next plane)
            int axis = -1, cell = 0, level = 0;
            if (it.count > 1) split_plane(g, it.lo, it.hi, axis, cell, level);
            if (it.count <= 1 || axis < 0) {
                // emit (an unsplittable part with count > 1 repeats its box: the reference count was fixed before the split)
                for (int k = 0; k < (it.count < 1 ? 1 : it.count); k++) {
                    out_lo[emitted] = f4(it.lo.x - pad, it.lo.y - pad, it.lo.z - pad, 0.0f);
                    out_hi[emitted] = f4(it.hi.x + pad, it.hi.y + pad, it.hi.z + pad, 0.0f);
                    emitted++;
                }
                break;
            }
            const float pos = split_comp(g.lo, axis) + (float)cell * split_comp(g.cell, axis);
            float3 l_lo, l_hi, r_lo, r_hi;
            split_boxes(v, axis, pos, it.lo, it.hi, l_lo, l_hi, r_lo, r_hi);
            const bool lv = split_box_valid(l_lo, l_hi), rv = split_box_valid(r_lo, r_hi);
            if (lv && rv) {
                const float wl = split_comp(l_hi, axis) - split_comp(l_lo, axis), wr = split_comp(r_hi, axis) - split_comp(r_lo, axis);
                int cl = (wl + wr) > 0.0f ? (int)((float)it.count * wl / (wl + wr) + 0.5f) : it.count / 2;
                cl = cl < 1 ? 1 : (cl > it.count - 1 ? it.count - 1 : cl);
                stack[sp++] = Item{r_lo, r_hi, it.count - cl};
                it = Item{l_lo, l_hi, cl};
            } else if (lv || rv) {
                const float3 nlo = lv ? l_lo : r_lo, nhi = lv ? l_hi : r_hi;
                // progress guard: the surviving side must not cross the same plane again (it cannot: it was cut at `pos`); should float
                // rounding leave the box unchanged, stop splitting it
                if (nlo.x == it.lo.x && nlo.y == it.lo.y && nlo.z == it.lo.z && nhi.x == it.hi.x && nhi.y == it.hi.y && nhi.z == it.hi.z) {
                    for (int k = 0; k < it.count; k++) {
                        out_lo[emitted] = f4(it.lo.x - pad, it.lo.y - pad, it.lo.z - pad, 0.0f);
                        out_hi[emitted] = f4(it.hi.x + pad, it.hi.y + pad, it.hi.z + pad, 0.0f);
                        emitted++;
                    }
                    break;
                }
                it.lo = nlo; it.hi = nhi;
            } else {
                // (numerically empty on both sides: keep the current box)
                for (int k = 0; k < it.count; k++) {
                    out_lo[emitted] = f4(it.lo.x - pad, it.lo.y - pad, it.lo.z - pad, 0.0f);
                    out_hi[emitted] = f4(it.hi.x + pad, it.hi.y + pad, it.hi.z + pad, 0.0f);
                    emitted++;
                }
                break;
            }
        }
    }
}

}  // namespace rfw
