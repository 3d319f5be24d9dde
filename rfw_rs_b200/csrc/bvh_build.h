// bvh_build.h — per-thread bodies of the device BVH builder (RFW_HD: compiled into the sm_100a
// kernels of builder.cu, and serially by the host logic harness in tests/hostemu).
//
// Replaces the reference's CPU builders: rtbvh BinnedSahBuilder + MBVH::construct
// (backends/gpu-rt/src/lib.rs:1345-1357 BLAS, :1576-1581 TLAS).
//
// Pipeline (one build = one mesh, or the TLAS over instance boxes):
//   1. prim boxes + centroid bounds            prim_bounds kernel (builder.cu)
//   2. 63-bit Morton keys of box centres       morton_body
//   3. radix sort (key, prim)                  radix_sort.cu
//   4. Karras 2012 binary radix tree           karras_body          (internal i in [0,n-2], leaf k = n-1+k)
//   5. bottom-up box fit + SAH forest costs    fit_cost_body        (Ylitie et al. 2017, eq. 1-5)
//   6. top-down collapse to 8-wide compressed  collapse_body        (one task per wide node, level by level)
#pragma once
#include "hd.h"

namespace rfw {

struct BuildParams {
    float c_node;  // cost of visiting one wide node
    float c_prim;  // cost of testing one primitive
    int pmax;      // max primitives per leaf child (<= 3: unary count in 3 meta bits)
    int treelet;   // binned-SAH refinement: LBVH subtrees of <= treelet primitives are kept, the tree above them is
                   // rebuilt top-down with binned SAH (builder.cu); 0 = plain LBVH
};

struct BuildArrays {
    int n;
    const float4* prim_lo;  // [n] boxes in submission order
    const float4* prim_hi;
    const uint64_t* keys;   // [n] sorted Morton keys
    const uint32_t* order;  // [n] sorted position -> prim index
    int* parent;            // [2n-1]
    int2* children;         // [n-1]
    int2* range;            // [n-1] first,last sorted positions covered
    float4* node_lo;        // [2n-1]
    float4* node_hi;
    float* cost;            // [(2n-1)*8]  cost[node*8 + (i-1)], i = 1..7;  [..+7] = surface area
    uint32_t* decision;     // [n-1]
    int* flags;             // [n-1] arrival counters for the bottom-up pass
};

struct CollapseOut {
    float4* nodes;           // 5 float4 per wide node
    uint32_t* leaf_prims;    // prim index per leaf slot (BLAS: triangle, TLAS: instance)
    uint32_t* node_counter;  // wide nodes allocated so far (starts at 1: the root)
    uint32_t* prim_counter;  // leaf slots allocated so far
};

// ---- 2. Morton ---------------------------------------------------------------------------------
RFW_HD uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1FFFFFu;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
RFW_HD uint64_t morton63(float3 c, float3 cmin, float3 cscale) {
    // cscale = 2^21 / extent (0 where the extent is 0)
    const float fx = fminf(fmaxf((c.x - cmin.x) * cscale.x, 0.0f), 2097151.0f);
    const float fy = fminf(fmaxf((c.y - cmin.y) * cscale.y, 0.0f), 2097151.0f);
    const float fz = fminf(fmaxf((c.z - cmin.z) * cscale.z, 0.0f), 2097151.0f);
    return (expand21((uint32_t)fx) << 2) | (expand21((uint32_t)fy) << 1) | expand21((uint32_t)fz);
}
RFW_HD void morton_body(int i, const float4* prim_lo, const float4* prim_hi, float3 cmin, float3 cscale, uint64_t* keys, uint32_t* vals) {
    const float3 c = (xyz(prim_lo[i]) + xyz(prim_hi[i])) * 0.5f;
    keys[i] = morton63(c, cmin, cscale);
    vals[i] = (uint32_t)i;
}

// ---- 4. Karras -----------------------------------------------------------------------------------
RFW_HD int karras_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);  // duplicate keys: fall back to the position
    return clz64(a ^ b);
}
RFW_HD void karras_body(int i, int n, const uint64_t* keys, int* parent, int2* children, int2* range) {
    const int d = (karras_delta(keys, n, i, i + 1) - karras_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = karras_delta(keys, n, i, i - d);
    int lmax = 2;
    while (karras_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (karras_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = karras_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    children[i] = make_int2(left, right);
    range[i] = make_int2(lo, hi);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

// Node ids: Karras internal i in [0, n-2]; leaf of sorted position k = n-1+k; nodes of the SAH-built top tree
// (builder.cu, optional) = 2n-1+j.  The per-inner-node arrays (children, range, decision, flags) are indexed by
// inner_index(): Karras internals first, top nodes behind them.
RFW_HD bool is_leaf_node(int node, int n) { return node >= n - 1 && node < 2 * n - 1; }
RFW_HD int inner_index(int node, int n) { return node < n - 1 ? node : node - n; }

// ---- 5. fit + SAH forest costs ---------------------------------------------------------------------
RFW_HD float box_area(float3 lo, float3 hi) {
    const float3 e = hi - lo;
    return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x);
}

// decision word: bit 0 = "leaf" for i = 1; bits 1..3 = split k of C_distribute(n, 8);
// bits 4+3(i-2) .. +2 = split k of C(n, i) for i = 2..7 (0 = "use C(n, i-1)")
RFW_HD uint32_t dec_leaf(uint32_t w) { return w & 1u; }
RFW_HD uint32_t dec_k8(uint32_t w) { return (w >> 1) & 7u; }
RFW_HD uint32_t dec_ki(uint32_t w, int i) { return (w >> (4 + 3 * (i - 2))) & 7u; }

RFW_HD void fit_cost_node(int node, const BuildArrays& A, const BuildParams& P) {
    const int2 ch = A.children[inner_index(node, A.n)];
    const float3 lo = min3(xyz(A.node_lo[ch.x]), xyz(A.node_lo[ch.y]));
    const float3 hi = max3(xyz(A.node_hi[ch.x]), xyz(A.node_hi[ch.y]));
    A.node_lo[node] = f4(lo.x, lo.y, lo.z, 0.0f);
    A.node_hi[node] = f4(hi.x, hi.y, hi.z, 0.0f);
    const float area = box_area(lo, hi);
    float cl[7], cr[7];
    for (int i = 0; i < 7; i++) { cl[i] = A.cost[(size_t)ch.x * 8 + i]; cr[i] = A.cost[(size_t)ch.y * 8 + i]; }
    float dist[9];
    uint32_t kd[9];
    for (int j = 2; j <= 8; j++) {
        float best = 3.0e38f; uint32_t bk = 1;
        for (int k = 1; k < j; k++) {
            if (k > 7 || j - k > 7) continue;
            const float c = cl[k - 1] + cr[j - k - 1];
            if (c < best) { best = c; bk = (uint32_t)k; }
        }
        dist[j] = best; kd[j] = bk;
    }
    const int2 rg = A.range[inner_index(node, A.n)];
    const int count = rg.y - rg.x + 1;
    // a node of the SAH-built top tree can never be a leaf: its primitives are not contiguous in the sorted order
    const float c_leaf = (count <= P.pmax && node < 2 * A.n - 1) ? area * (float)count * P.c_prim : 3.0e38f;
    const float c_internal = dist[8] + area * P.c_node;
    uint32_t w = (c_leaf <= c_internal) ? 1u : 0u;
    w |= kd[8] << 1;
    float c[8];
    c[1] = fminf(c_leaf, c_internal);
    for (int i = 2; i <= 7; i++) {
        if (dist[i] < c[i - 1]) { c[i] = dist[i]; w |= kd[i] << (4 + 3 * (i - 2)); }
        else c[i] = c[i - 1];
    }
    for (int i = 1; i <= 7; i++) A.cost[(size_t)node * 8 + (i - 1)] = c[i];
    A.cost[(size_t)node * 8 + 7] = area;
    A.decision[inner_index(node, A.n)] = w;
}

// the leaf part of the bottom-up pass alone (box, costs); the climb is the caller's
RFW_HD void fit_cost_leaf(int k, const BuildArrays& A, const BuildParams& P) {
    const int leaf = A.n - 1 + k;
    const uint32_t prim = A.order[k];
    const float4 lo = A.prim_lo[prim], hi = A.prim_hi[prim];
    A.node_lo[leaf] = lo;
    A.node_hi[leaf] = hi;
    const float area = box_area(xyz(lo), xyz(hi));
    for (int i = 0; i < 7; i++) A.cost[(size_t)leaf * 8 + i] = area * P.c_prim;
    A.cost[(size_t)leaf * 8 + 7] = area;
}

// one thread per sorted leaf k: write the leaf, then climb; the second thread to arrive at a node computes it
template <bool CTA_SCOPE = false>
RFW_HD void fit_cost_body(int k, const BuildArrays& A, const BuildParams& P) {
    const int n = A.n;
    const int leaf = n - 1 + k;
    const uint32_t prim = A.order[k];
    const float4 lo = A.prim_lo[prim], hi = A.prim_hi[prim];
    A.node_lo[leaf] = lo;
    A.node_hi[leaf] = hi;
    const float area = box_area(xyz(lo), xyz(hi));
    for (int i = 0; i < 7; i++) A.cost[(size_t)leaf * 8 + i] = area * P.c_prim;
    A.cost[(size_t)leaf * 8 + 7] = area;
    if (n == 1) return;
    int cur = A.parent[leaf];
    while (cur >= 0) {
        thread_fence_scope<CTA_SCOPE>();
        const int old = atomic_add(&A.flags[inner_index(cur, n)], 1);
        if (old == 0) return;  // first arrival: the sibling subtree is not finished yet
        thread_fence_scope<CTA_SCOPE>();
        fit_cost_node(cur, A, P);
        cur = A.parent[cur];
    }
}

// ---- 6. collapse -------------------------------------------------------------------------------
struct WideChild {
    int bnode;     // binary node id
    int first;     // leaf: first sorted position
    int count;     // leaf: primitive count; 0 = inner child
};

// children of the wide node rooted at binary node `bnode` (forced inner), following the DP decisions
RFW_HD int gather_wide_children(int bnode, const BuildArrays& A, WideChild* out) {
    const int n = A.n;
    int nch = 0;
    if (is_leaf_node(bnode, n)) {  // single-primitive tree
        out[0].bnode = bnode; out[0].first = bnode - (n - 1); out[0].count = 1;
        return 1;
    }
    int sn[16], sb[16];
    int sp = 0;
    {
        const uint32_t w = A.decision[inner_index(bnode, n)];
        const int k = (int)dec_k8(w);
        const int2 ch = A.children[inner_index(bnode, n)];
        sn[sp] = ch.y; sb[sp] = 8 - k; sp++;
        sn[sp] = ch.x; sb[sp] = k; sp++;
    }
    while (sp > 0) {
        sp--;
        const int m = sn[sp];
        int j = sb[sp];
        if (is_leaf_node(m, n)) {
            out[nch].bnode = m; out[nch].first = m - (n - 1); out[nch].count = 1; nch++;
            continue;
        }
        const uint32_t w = A.decision[inner_index(m, n)];
        while (j > 1 && dec_ki(w, j) == 0u) j--;  // "use C(m, j-1)"
        if (j == 1) {
            if (dec_leaf(w)) {
                const int2 rg = A.range[inner_index(m, n)];
                out[nch].bnode = m; out[nch].first = rg.x; out[nch].count = rg.y - rg.x + 1; nch++;
            } else {
                out[nch].bnode = m; out[nch].first = 0; out[nch].count = 0; nch++;
            }
            continue;
        }
        const int k = (int)dec_ki(w, j);
        const int2 ch = A.children[inner_index(m, n)];
        sn[sp] = ch.y; sb[sp] = j - k; sp++;
        sn[sp] = ch.x; sb[sp] = k; sp++;
    }
    return nch;
}

RFW_HD uint32_t exponent_for_extent(float extent) {
    // smallest biased exponent e8 with 2^(e8-127) * 255 >= extent
    if (!(extent > 0.0f)) return 0u;
    int e;
    frexpf(extent / 255.0f, &e);  // extent/255 = m * 2^e, m in [0.5,1)  =>  2^e > extent/255
    while (ldexp(255.0, e - 1) >= (double)extent) e--;
    while (ldexp(255.0, e) < (double)extent) e++;
    int e8 = e + 127;
    if (e8 < 1) e8 = 1;
    if (e8 > 254) e8 = 254;
    return (uint32_t)e8;
}

// One wide node: gather children, assign octant slots, allocate child / leaf ranges, quantise, write.
// Inner children are appended to the next level's task queue.
RFW_HD void collapse_body(int2 task, const BuildArrays& A, const CollapseOut& O, int2* next_queue, uint32_t* next_count) {
    const int bnode = task.x;
    const uint32_t widx = (uint32_t)task.y;
    WideChild ch[8];
    const int nch = gather_wide_children(bnode, A, ch);
    const float3 plo = xyz(A.node_lo[bnode]), phi = xyz(A.node_hi[bnode]);
    const float3 pc = (plo + phi) * 0.5f;

    // greedy octant slot assignment: slot bit 2/1/0 set <=> child lies towards +x/+y/+z of the parent centre
    float score[8][8];
    for (int i = 0; i < nch; i++) {
        const float3 c = (xyz(A.node_lo[ch[i].bnode]) + xyz(A.node_hi[ch[i].bnode])) * 0.5f - pc;
        for (int s = 0; s < 8; s++) score[i][s] = ((s & 4) ? c.x : -c.x) + ((s & 2) ? c.y : -c.y) + ((s & 1) ? c.z : -c.z);
    }
    int slot_child[8];
    for (int s = 0; s < 8; s++) slot_child[s] = -1;
    uint32_t assigned = 0;
    for (int it = 0; it < nch; it++) {
        float best = -3.0e38f; int bi = -1, bs = -1;
        for (int i = 0; i < nch; i++) {
            if (assigned & (1u << i)) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_child[s] >= 0) continue;
                if (score[i][s] > best) { best = score[i][s]; bi = i; bs = s; }
            }
        }
        slot_child[bs] = bi;
        assigned |= 1u << bi;
    }

    int n_inner = 0, n_leaf_prims = 0;
    for (int i = 0; i < nch; i++) {
        if (ch[i].count == 0) n_inner++;
        else n_leaf_prims += ch[i].count;
    }
    const uint32_t child_base = n_inner ? atomic_add(O.node_counter, (uint32_t)n_inner) : 0u;
    const uint32_t prim_base = n_leaf_prims ? atomic_add(O.prim_counter, (uint32_t)n_leaf_prims) : 0u;

    const uint32_t ex = exponent_for_extent(phi.x - plo.x), ey = exponent_for_extent(phi.y - plo.y), ez = exponent_for_extent(phi.z - plo.z);
    const double isx = ldexp(1.0, 127 - (int)ex), isy = ldexp(1.0, 127 - (int)ey), isz = ldexp(1.0, 127 - (int)ez);

    uint32_t imask = 0;
    uint32_t meta[2] = {0, 0}, qlx[2] = {0, 0}, qly[2] = {0, 0}, qlz[2] = {0, 0}, qhx[2] = {0, 0}, qhy[2] = {0, 0}, qhz[2] = {0, 0};
    uint32_t inner_rank = 0, prim_off = 0;
    uint32_t queue_base = 0;
    if (n_inner) queue_base = atomic_add(next_count, (uint32_t)n_inner);
    for (int s = 0; s < 8; s++) {
        const int half = s >> 2, sh = 8 * (s & 3);
        const int i = slot_child[s];
        if (i < 0) {  // empty slot: inverted box, never hit
            qlx[half] |= 255u << sh; qly[half] |= 255u << sh; qlz[half] |= 255u << sh;
            continue;
        }
        uint32_t m;
        if (ch[i].count == 0) {
            imask |= 1u << s;
            m = (1u << 5) | (24u + (uint32_t)s);
            next_queue[queue_base + inner_rank] = make_int2(ch[i].bnode, (int)(child_base + inner_rank));
            inner_rank++;
        } else {
            const uint32_t unary = (1u << ch[i].count) - 1u;  // 1, 3, 7
            m = (unary << 5) | prim_off;
            for (int j = 0; j < ch[i].count; j++) O.leaf_prims[prim_base + prim_off + j] = A.order[ch[i].first + j];
            prim_off += (uint32_t)ch[i].count;
        }
        meta[half] |= m << sh;
        const float3 clo = xyz(A.node_lo[ch[i].bnode]), chi = xyz(A.node_hi[ch[i].bnode]);
        double q;
        q = floor(((double)clo.x - (double)plo.x) * isx); qlx[half] |= (uint32_t)(q < 0 ? 0 : (q > 255 ? 255 : q)) << sh;
        q = floor(((double)clo.y - (double)plo.y) * isy); qly[half] |= (uint32_t)(q < 0 ? 0 : (q > 255 ? 255 : q)) << sh;
        q = floor(((double)clo.z - (double)plo.z) * isz); qlz[half] |= (uint32_t)(q < 0 ? 0 : (q > 255 ? 255 : q)) << sh;
        q = ceil(((double)chi.x - (double)plo.x) * isx); qhx[half] |= (uint32_t)(q < 0 ? 0 : (q > 255 ? 255 : q)) << sh;
        q = ceil(((double)chi.y - (double)plo.y) * isy); qhy[half] |= (uint32_t)(q < 0 ? 0 : (q > 255 ? 255 : q)) << sh;
        q = ceil(((double)chi.z - (double)plo.z) * isz); qhz[half] |= (uint32_t)(q < 0 ? 0 : (q > 255 ? 255 : q)) << sh;
    }
    float4* out = O.nodes + (size_t)widx * NODE_F4;
    if (NODE_F4 > 5) out[5] = f4(0.0f, 0.0f, 0.0f, 0.0f);  // pad of the 32-B aligned stride (hashed by the BVH checksum)
    out[0] = f4(plo.x, plo.y, plo.z, u2f(ex | (ey << 8) | (ez << 16) | (imask << 24)));
    out[1] = f4(u2f(child_base), u2f(prim_base), u2f(meta[0]), u2f(meta[1]));
    out[2] = f4(u2f(qlx[0]), u2f(qlx[1]), u2f(qly[0]), u2f(qly[1]));
    out[3] = f4(u2f(qlz[0]), u2f(qlz[1]), u2f(qhx[0]), u2f(qhx[1]));
    out[4] = f4(u2f(qhy[0]), u2f(qhy[1]), u2f(qhz[0]), u2f(qhz[1]));
}

}  // namespace rfw
