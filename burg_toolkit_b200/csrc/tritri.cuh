// tritri.cuh -- FCL's exact triangle/triangle intersection test (float64, 17 separating axes), shared by the
// gripper-vs-scene kernel (collide.cu) and the mesh-vs-mesh kernel (scene.cu).
#pragma once

#include "common.cuh"

namespace bt {

// FCL Intersect::project6: true if `ax` does NOT separate {p1,p2,p3} from {q1,q2,q3}
__device__ __forceinline__ bool project6(d3 ax, d3 p1, d3 p2, d3 p3, d3 q1, d3 q2, d3 q3) {
    const double P1 = dot3(ax, p1), P2 = dot3(ax, p2), P3 = dot3(ax, p3);
    const double Q1 = dot3(ax, q1), Q2 = dot3(ax, q2), Q3 = dot3(ax, q3);
    const double mn1 = fmin(P1, fmin(P2, P3));
    const double mx2 = fmax(Q1, fmax(Q2, Q3));
    if (mn1 > mx2) return false;
    const double mx1 = fmax(P1, fmax(P2, P3));
    const double mn2 = fmin(Q1, fmin(Q2, Q3));
    if (mn2 > mx1) return false;
    return true;
}

// FCL Intersect::intersect_Triangle(P1,P2,P3, Q1,Q2,Q3): P = scene triangle, Q = posed gripper triangle
__device__ inline bool tri_tri_intersect(d3 P1, d3 P2, d3 P3, d3 Q1, d3 Q2, d3 Q3) {
    const d3 p1 = sub3(P1, P1), p2 = sub3(P2, P1), p3 = sub3(P3, P1);
    const d3 q1 = sub3(Q1, P1), q2 = sub3(Q2, P1), q3 = sub3(Q3, P1);
    const d3 e1 = sub3(p2, p1), e2 = sub3(p3, p2);
    const d3 n1 = cross3(e1, e2);
    if (!project6(n1, p1, p2, p3, q1, q2, q3)) return false;
    const d3 f1 = sub3(q2, q1), f2 = sub3(q3, q2);
    const d3 m1 = cross3(f1, f2);
    if (!project6(m1, p1, p2, p3, q1, q2, q3)) return false;
    const d3 e3 = sub3(p1, p3), f3 = sub3(q1, q3);
    if (!project6(cross3(e1, f1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e1, f2), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e1, f3), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e2, f1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e2, f2), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e2, f3), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e3, f1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e3, f2), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e3, f3), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e1, n1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e2, n1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(e3, n1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(f1, m1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(f2, m1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross3(f3, m1), p1, p2, p3, q1, q2, q3)) return false;
    return true;
}

}  // namespace bt
