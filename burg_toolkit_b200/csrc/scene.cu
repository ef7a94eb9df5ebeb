// scene.cu -- mesh-vs-mesh collision for scene composition.
//
// Reference being replaced: mesh_processing.collisions (burg_toolkit/mesh_processing.py:266-281), used by
// Scene.colliding_instances (core.py:666-687) and sampling.sample_scene (sampling.py:498-569): a trimesh
// CollisionManager.in_collision_internal -> FCL mesh/mesh query per pair of objects.  FCL's answer is "does any triangle
// of one mesh intersect any triangle of the other" (surface test: containment is not a collision), decided by the
// 17-axis separating-axis test of Intersect::intersect_Triangle.
//
// Here: both meshes are bt_mesh objects in WORLD coordinates (the reference hands over posed copies, core.py:266-276).
// One thread per triangle of the smaller mesh walks the other mesh's LBVH with the triangle's float32 box (rounded
// outward, so the broad phase is result-neutral); at the leaves the float64 17-axis test decides.  The first hit raises
// the pair's flag and every other thread of the launch stops at its next flag check.
#include <math.h>

#include "common.cuh"
#include "tritri.cuh"

namespace bt {

constexpr int MSTACK = 64;

// walk `nodes` (tree of mesh T) with every triangle of mesh Q; tree_is_first: T's triangles play FCL's P role
__global__ void __launch_bounds__(128)
mesh_pair_kernel(const float4* __restrict__ nodes, const double* __restrict__ tree_tris, const double* __restrict__ query_tris,
                 int64_t n_query, int tree_is_first, int* flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_query) return;
    if (*reinterpret_cast<volatile int*>(flag)) return;
    const double* qv = query_tris + 9 * i;
    const d3 Q1 = ld3(qv), Q2 = ld3(qv + 3), Q3 = ld3(qv + 6);
    // float32 box of the query triangle, rounded outward
    const float x0 = __double2float_rd(fmin(Q1.x, fmin(Q2.x, Q3.x))), x1 = __double2float_ru(fmax(Q1.x, fmax(Q2.x, Q3.x)));
    const float y0 = __double2float_rd(fmin(Q1.y, fmin(Q2.y, Q3.y))), y1 = __double2float_ru(fmax(Q1.y, fmax(Q2.y, Q3.y)));
    const float z0 = __double2float_rd(fmin(Q1.z, fmin(Q2.z, Q3.z))), z1 = __double2float_ru(fmax(Q1.z, fmax(Q2.z, Q3.z)));
    int stack[MSTACK];
    int sp = 0, cur = 0, steps = 0;
    while (true) {
        if (cur >= 0) {
            float4 q0, q1, q2, q3;                                     // 64-byte node: 2 x LDG.256
            ldg256(nodes + 4 * (int64_t)cur, q0, q1);
            ldg256(nodes + 4 * (int64_t)cur + 2, q2, q3);
            // left box (q0.x q0.y q0.z)-(q0.w q1.x q1.y), right box (q1.z q1.w q2.x)-(q2.y q2.z q2.w)
            const bool hl = (q0.x <= x1) & (q0.w >= x0) & (q0.y <= y1) & (q1.x >= y0) & (q0.z <= z1) & (q1.y >= z0);
            const bool hr = (q1.z <= x1) & (q2.y >= x0) & (q1.w <= y1) & (q2.z >= y0) & (q2.x <= z1) & (q2.w >= z0);
            const int l = __float_as_int(q3.x), r = __float_as_int(q3.y);
            if (hl && hr) { stack[sp++] = r; cur = l; }
            else if (hl) cur = l;
            else if (hr) cur = r;
            else { if (sp == 0) return; cur = stack[--sp]; }
            if ((++steps & 63) == 0 && *reinterpret_cast<volatile int*>(flag)) return;
        } else {
            const double* tv = tree_tris + 9 * (int64_t)(~cur);
            const d3 T1 = ld3(tv), T2 = ld3(tv + 3), T3 = ld3(tv + 6);
            const bool hit = tree_is_first ? tri_tri_intersect(T1, T2, T3, Q1, Q2, Q3) : tri_tri_intersect(Q1, Q2, Q3, T1, T2, T3);
            if (hit) { atomicExch(flag, 1); return; }
            if (sp == 0) return;
            cur = stack[--sp];
        }
    }
}

}  // namespace bt

using namespace bt;

extern "C" int bt_mesh_pair_collide(const bt_mesh* a, const bt_mesh* b, int32_t* d_flag, void* stream) {
    BT_REQUIRE(a && b && d_flag, "null pointer");
    BT_REQUIRE(a->m.device == b->m.device, "the two meshes live on different devices");
    BT_REQUIRE(a->m.info.max_depth + 2 <= 64 && b->m.info.max_depth + 2 <= 64, "LBVH too deep for the traversal stack (64)");
    cudaStream_t st = (cudaStream_t)stream;
    BT_CUDA(cudaSetDevice(a->m.device));
    BT_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), st));
    if (a->m.nF == 0 || b->m.nF == 0) return BT_OK;
    // bounding boxes first (result-neutral: intersecting triangles have overlapping boxes)
    // (info.bounds_* are the float64 bounds rounded to nearest float32: widen each by an ulp or two before comparing)
    auto below = [](float v) { return v - (4e-7f * fabsf(v) + 1e-30f); };
    auto above = [](float v) { return v + (4e-7f * fabsf(v) + 1e-30f); };
    for (int k = 0; k < 3; ++k)
        if (below(a->m.info.bounds_min[k]) > above(b->m.info.bounds_max[k]) ||
            below(b->m.info.bounds_min[k]) > above(a->m.info.bounds_max[k])) return BT_OK;
    // the smaller mesh queries the tree of the larger one; FCL's P role stays with `a`
    const bool tree_is_a = a->m.nF >= b->m.nF;
    const MeshDev& tree = tree_is_a ? a->m : b->m;
    const MeshDev& qry = tree_is_a ? b->m : a->m;
    const unsigned grid = (unsigned)ceil_div(qry.nF, 128);
    mesh_pair_kernel<<<grid, 128, 0, st>>>(tree.nodes, tree.tri_verts, qry.tri_verts, qry.nF, tree_is_a ? 1 : 0, d_flag);
    BT_LAUNCH_CHECK();
    return BT_OK;
}
