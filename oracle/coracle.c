/*
 * TEST INFRASTRUCTURE (not product code) -- plain-C port of the reference's AGS hot path.
 *
 * Restates, for sizes numpy cannot finish in seconds, exactly what oracle/port.py + oracle/leaves.py state
 * (and those are pinned against the reference's own Python source, tests/golden/):
 *   - trimesh ray_triangle_id narrow phase, all crossings        (reference sampling.py:231-232)
 *   - origin / width / sign / cone / z filters per hit           (sampling.py:239-319)
 *   - compute_interpolated_vertex_normals                         (mesh_processing.py:110-153)
 *   - construct_grasp_set                                         (sampling.py:132-183)
 *   - check_collisions = any triangle pair intersecting, FCL 17-axis test (sampling.py:354-397)
 * A median-split AABB tree stands in for trimesh's rtree / FCL's BVH (result-neutral broad phases); OpenMP
 * spreads reference points / grasps over the host cores.  Compiled with -ffp-contract=off so every multiply
 * and add rounds separately, like numpy.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HIT_ORIGIN 1u
#define HIT_WIDTH 2u
#define HIT_SIGN 4u
#define HIT_CONE 8u
#define HIT_BELOW_Z 16u

typedef struct { double lo[3], hi[3]; int left, right, start, count; } Node;

typedef struct {
    int nV, nF, n_nodes;
    double *V, *VN;      /* [nV,3] */
    int *F;              /* [nF,3] */
    double *fn;          /* [nF,3] face normals (trimesh unitize) */
    int *order;          /* leaf triangle order */
    Node *nodes;
} CoMesh;

typedef struct {         /* mirrors struct bt_hit of include/burg_b200.h */
    double loc[3], nrm[3], t, angle;
    int32_t tri;
    uint32_t flags;
} CoHit;

static double dot3(const double *a, const double *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double *a) { return sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }

/* ---- tree ---------------------------------------------------------------------------------------------- */
static void tri_box(const CoMesh *m, int f, double *lo, double *hi) {
    for (int k = 0; k < 3; ++k) { lo[k] = 1e300; hi[k] = -1e300; }
    for (int v = 0; v < 3; ++v) {
        const double *p = m->V + 3 * m->F[3 * f + v];
        for (int k = 0; k < 3; ++k) { if (p[k] < lo[k]) lo[k] = p[k]; if (p[k] > hi[k]) hi[k] = p[k]; }
    }
}

static int g_axis;
static const double *g_cent;
static int cmp_cent(const void *a, const void *b) {
    double x = g_cent[3 * (*(const int *)a) + g_axis], y = g_cent[3 * (*(const int *)b) + g_axis];
    return (x > y) - (x < y);
}

static int build_node(CoMesh *m, const double *cent, int start, int count, double pad) {
    int id = m->n_nodes++;
    Node *n = m->nodes + id;
    double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (int k = 0; k < 3; ++k) { n->lo[k] = 1e300; n->hi[k] = -1e300; }
    for (int i = start; i < start + count; ++i) {
        double lo[3], hi[3];
        tri_box(m, m->order[i], lo, hi);
        for (int k = 0; k < 3; ++k) {
            if (lo[k] - pad < n->lo[k]) n->lo[k] = lo[k] - pad;
            if (hi[k] + pad > n->hi[k]) n->hi[k] = hi[k] + pad;
            double c = cent[3 * m->order[i] + k];
            if (c < clo[k]) clo[k] = c;
            if (c > chi[k]) chi[k] = c;
        }
    }
    n->start = start; n->count = count; n->left = n->right = -1;
    if (count <= 4) return id;
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    g_axis = axis; g_cent = cent;
    qsort(m->order + start, count, sizeof(int), cmp_cent);
    int half = count / 2;
    int l = build_node(m, cent, start, half, pad);
    int r = build_node(m, cent, start + half, count - half, pad);
    n = m->nodes + id;            /* (nodes is pre-allocated, pointer is stable; re-read for clarity) */
    n->left = l; n->right = r; n->count = 0;
    return id;
}

CoMesh *co_mesh_create(const double *V, int nV, const int *F, int nF, const double *VN) {
    CoMesh *m = (CoMesh *)calloc(1, sizeof(CoMesh));
    m->nV = nV; m->nF = nF;
    m->V = (double *)malloc(sizeof(double) * 3 * nV); memcpy(m->V, V, sizeof(double) * 3 * nV);
    m->F = (int *)malloc(sizeof(int) * 3 * nF); memcpy(m->F, F, sizeof(int) * 3 * nF);
    m->VN = (double *)calloc(3 * (size_t)nV, sizeof(double));
    m->fn = (double *)malloc(sizeof(double) * 3 * nF);
    double *cr = (double *)malloc(sizeof(double) * 3 * nF);
    double *cent = (double *)malloc(sizeof(double) * 3 * nF);
    double max_abs = 0;
    for (int i = 0; i < 3 * nV; ++i) if (fabs(V[i]) > max_abs) max_abs = fabs(V[i]);
    for (int f = 0; f < nF; ++f) {
        const double *a = V + 3 * F[3 * f], *b = V + 3 * F[3 * f + 1], *c = V + 3 * F[3 * f + 2];
        double e0[3], e1[3];
        for (int k = 0; k < 3; ++k) { e0[k] = b[k] - a[k]; e1[k] = c[k] - a[k]; cent[3 * f + k] = (a[k] + b[k] + c[k]) / 3.0; }
        cross3(e0, e1, cr + 3 * f);
        double nn = norm3(cr + 3 * f);
        if (nn > 1e-13) { double inv = 1.0 / nn; for (int k = 0; k < 3; ++k) m->fn[3 * f + k] = cr[3 * f + k] * inv; }
        else for (int k = 0; k < 3; ++k) m->fn[3 * f + k] = 0.0;
    }
    if (VN) memcpy(m->VN, VN, sizeof(double) * 3 * nV);
    else {
        for (int v = 0; v < 3; ++v)
            for (int f = 0; f < nF; ++f)
                for (int k = 0; k < 3; ++k) m->VN[3 * F[3 * f + v] + k] += cr[3 * f + k];
        for (int v = 0; v < nV; ++v) {
            double nn = norm3(m->VN + 3 * v);
            if (nn > 0) for (int k = 0; k < 3; ++k) m->VN[3 * v + k] /= nn;
            else { m->VN[3 * v] = 0; m->VN[3 * v + 1] = 0; m->VN[3 * v + 2] = 1; }
        }
    }
    m->order = (int *)malloc(sizeof(int) * nF);
    for (int f = 0; f < nF; ++f) m->order[f] = f;
    m->nodes = (Node *)malloc(sizeof(Node) * (2 * (size_t)nF + 1));
    m->n_nodes = 0;
    build_node(m, cent, 0, nF, 1e-6 * max_abs + 1e-12);
    free(cr); free(cent);
    return m;
}

void co_mesh_free(CoMesh *m) {
    if (!m) return;
    free(m->V); free(m->VN); free(m->F); free(m->fn); free(m->order); free(m->nodes); free(m);
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the benchmark's CPU arm asks for all host threads explicitly */
void co_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int co_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- ray casting ------------------------------------------------------------------------------------------ */
static int ray_box(const Node *n, const double *o, const double *inv) {
    double tn = -1e300, tf = 1e300;
    for (int k = 0; k < 3; ++k) {
        double t0 = (n->lo[k] - o[k]) * inv[k], t1 = (n->hi[k] - o[k]) * inv[k];
        if (t0 != t0 || t1 != t1) continue;                 /* 0 * inf: origin on the slab boundary */
        if (t0 > t1) { double s = t0; t0 = t1; t1 = s; }
        if (t0 > tn) tn = t0;
        if (t1 < tf) tf = t1;
    }
    return tn <= tf && tf >= -1e-5;
}

/* trimesh ray_triangle_id narrow phase for one pair; returns 1 and *t on a hit */
static int ray_tri(const CoMesh *m, int f, const double *o, const double *d, double *t_out) {
    const double *v0 = m->V + 3 * m->F[3 * f], *v1 = m->V + 3 * m->F[3 * f + 1], *v2 = m->V + 3 * m->F[3 * f + 2];
    const double *n = m->fn + 3 * f;
    double ov[3] = {v0[0] - o[0], v0[1] - o[1], v0[2] - o[2]};
    double proj_ori = dot3(ov, n), proj_dir = dot3(d, n);
    if (!(fabs(proj_dir) > 1e-5)) return 0;
    double t = proj_ori / proj_dir;
    double p[3], vec[3], w[3], e0[3], e1[3];
    for (int k = 0; k < 3; ++k) {
        p[k] = d[k] * t; p[k] += o[k];
        e0[k] = v1[k] - v0[k]; e1[k] = v2[k] - v0[k];
        w[k] = p[k] - v0[k];
        vec[k] = p[k] - o[k];
    }
    double d00 = dot3(e0, e0), d01 = dot3(e0, e1), d02 = dot3(e0, w), d11 = dot3(e1, e1), d12 = dot3(e1, w);
    double inv = 1.0 / (d00 * d11 - d01 * d01);
    double b2 = (d00 * d12 - d01 * d02) * inv;
    double b1 = (d11 * d02 - d01 * d12) * inv;
    double b0 = 1 - b1 - b2;
    const double lo = -1e-13, hi = 1 + 1e-13;
    if (!(b0 > lo && b1 > lo && b2 > lo && b0 < hi && b1 < hi && b2 < hi)) return 0;
    if (!(dot3(vec, d) > -1e-6)) return 0;
    *t_out = t;
    return 1;
}

typedef struct { double t; int tri; } THit;
static int cmp_thit(const void *a, const void *b) {
    const THit *x = (const THit *)a, *y = (const THit *)b;
    if (x->t != y->t) return (x->t > y->t) - (x->t < y->t);
    return (x->tri > y->tri) - (x->tri < y->tri);
}

typedef struct {
    double angle_max, opening_width, width_tolerance, min_grasp_width, no_contact_below_z;
    int32_t use_below_z, max_hits_per_ray;
} CoParams;

static void interp_normal(const CoMesh *m, int tri, const double *p, double *out) {
    double w[3], s[3] = {0, 0, 0};
    int onehot = -1;
    for (int k = 0; k < 3; ++k) {
        const double *v = m->V + 3 * m->F[3 * tri + k];
        double dv[3] = {v[0] - p[0], v[1] - p[1], v[2] - p[2]};
        double dist = norm3(dv);
        w[k] = 1.0 / sqrt(dist);
        if (dist == 0.0) onehot = k;
    }
    if (onehot >= 0) for (int k = 0; k < 3; ++k) w[k] = (k == onehot) ? 1.0 : 0.0;
    const double *n0 = m->VN + 3 * m->F[3 * tri], *n1 = m->VN + 3 * m->F[3 * tri + 1], *n2 = m->VN + 3 * m->F[3 * tri + 2];
    for (int c = 0; c < 3; ++c) s[c] = ((w[0] * n0[c] + w[1] * n1[c]) + w[2] * n2[c]) / 3.0;
    double nn = norm3(s);
    for (int c = 0; c < 3; ++c) out[c] = s[c] / nn;
}

/* all hits of one ray, canonical order + 1e-8 de-duplication + per-hit filters; returns the count (<= cap) */
static int cast_one(const CoMesh *m, const double *o, const double *d, const CoParams *P, CoHit *out, int cap) {
    double inv[3];
    for (int k = 0; k < 3; ++k) inv[k] = 1.0 / d[k];
    THit th[256];
    int nh = 0;
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Node *n = m->nodes + stack[--sp];
        if (!ray_box(n, o, inv)) continue;
        if (n->count > 0) {
            for (int i = n->start; i < n->start + n->count; ++i) {
                double t;
                if (nh < 256 && ray_tri(m, m->order[i], o, d, &t)) { th[nh].t = t; th[nh].tri = m->order[i]; ++nh; }
            }
        } else { stack[sp++] = n->left; stack[sp++] = n->right; }
    }
    qsort(th, nh, sizeof(THit), cmp_thit);
    int n_out = 0;
    for (int i = 0; i < nh && n_out < cap; ++i) {
        double p[3];
        for (int k = 0; k < 3; ++k) { p[k] = d[k] * th[i].t; p[k] += o[k]; }
        int dup = 0;
        for (int j = 0; j < n_out && !dup; ++j) {
            dup = 1;
            for (int k = 0; k < 3; ++k)
                if (rint(out[j].loc[k] * 1e8 - 1e-6) != rint(p[k] * 1e8 - 1e-6)) dup = 0;
        }
        if (dup) continue;
        CoHit *h = out + n_out++;
        /* masks in the reference's order; flags = the first one that rejects the hit, later quantities stay 0 / NaN */
        uint32_t flags = 0;
        int close = 1;
        double dv[3], ang = NAN;
        for (int k = 0; k < 3; ++k) {
            if (!(fabs(p[k] - o[k]) <= 1e-11 + 1e-5 * fabs(o[k]))) close = 0;
            dv[k] = p[k] - o[k];
            h->loc[k] = p[k];
            h->nrm[k] = 0.0;
        }
        double dist = norm3(dv);
        if (close) flags = HIT_ORIGIN;
        else if (!((dist <= P->opening_width - P->width_tolerance) || (dist >= P->min_grasp_width))) flags = HIT_WIDTH;
        else {
            interp_normal(m, th[i].tri, p, h->nrm);
            double dotp = dot3(dv, h->nrm), cr[3];
            cross3(dv, h->nrm, cr);
            ang = atan(norm3(cr) / dotp);
            if (!(dotp > 0.0)) flags = HIT_SIGN;
            else if (!(ang <= P->angle_max)) flags = HIT_CONE;
            else if (P->use_below_z && !(p[2] > P->no_contact_below_z)) flags = HIT_BELOW_Z;
        }
        h->t = th[i].t; h->angle = ang; h->tri = th[i].tri; h->flags = flags;
    }
    return n_out;
}

/* same outputs as bt_ags_cast (host memory): hit_count [n_ref*n_rays], hits [n_ref*n_rays*cap] */
void co_ags_cast(const CoMesh *m, const double *ref_pts, const double *dirs, int64_t n_ref, int n_rays,
                 const CoParams *P, int32_t *hit_count, CoHit *hits) {
    const int cap = P->max_hits_per_ray;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ray = 0; ray < n_ref * n_rays; ++ray)
        hit_count[ray] = cast_one(m, ref_pts + 3 * (ray / n_rays), dirs + 3 * ray, P, hits + ray * cap, cap);
}

/* ---- orientation expansion (k = 1 target per reference point) --------------------------------------------- */
static void expand_one(const double *p_r, const double *q, const double *u, int n_o, const double *cs, const double *sn,
                       float *rows, double *contacts) {
    double d[3], c[3], x[3], y0[3], z0[3];
    for (int k = 0; k < 3; ++k) { d[k] = q[k] - p_r[k]; c[k] = p_r[k] + 0.5 * d[k]; }
    double dist = norm3(d);
    for (int k = 0; k < 3; ++k) x[k] = d[k] / dist;
    cross3(x, u, y0);
    double ny = norm3(y0);
    for (int k = 0; k < 3; ++k) y0[k] /= ny;
    cross3(x, y0, z0);
    double nz = norm3(z0);
    for (int k = 0; k < 3; ++k) z0[k] /= nz;
    for (int o = 0; o < n_o; ++o) {
        float *g = rows + 14 * o;
        for (int k = 0; k < 3; ++k) {
            g[k] = (float)c[k];
            g[3 + 3 * k] = (float)x[k];
            g[4 + 3 * k] = (float)fma(z0[k], sn[o], y0[k] * cs[o]);
            g[5 + 3 * k] = (float)fma(z0[k], cs[o], y0[k] * (-sn[o]));
        }
        g[12] = 0.f; g[13] = (float)dist;
        if (contacts) { memcpy(contacts + 6 * o, p_r, 24); memcpy(contacts + 6 * o + 3, q, 24); }
    }
}

/* ---- collision ---------------------------------------------------------------------------------------------- */
static int project6(const double *ax, const double (*p)[3], const double (*q)[3]) {
    double P1 = dot3(ax, p[0]), P2 = dot3(ax, p[1]), P3 = dot3(ax, p[2]);
    double Q1 = dot3(ax, q[0]), Q2 = dot3(ax, q[1]), Q3 = dot3(ax, q[2]);
    double mn1 = fmin(P1, fmin(P2, P3)), mx2 = fmax(Q1, fmax(Q2, Q3));
    if (mn1 > mx2) return 0;
    double mx1 = fmax(P1, fmax(P2, P3)), mn2 = fmin(Q1, fmin(Q2, Q3));
    if (mn2 > mx1) return 0;
    return 1;
}

/* FCL Intersect::intersect_Triangle: P = scene triangle, Q = posed gripper triangle */
static int tri_tri(const double *P, const double *Q) {
    double p[3][3], q[3][3], e[3][3], f[3][3], n1[3], m1[3], ax[3];
    for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) { p[v][k] = P[3 * v + k] - P[k]; q[v][k] = Q[3 * v + k] - P[k]; }
    for (int k = 0; k < 3; ++k) {
        e[0][k] = p[1][k] - p[0][k]; e[1][k] = p[2][k] - p[1][k]; e[2][k] = p[0][k] - p[2][k];
        f[0][k] = q[1][k] - q[0][k]; f[1][k] = q[2][k] - q[1][k]; f[2][k] = q[0][k] - q[2][k];
    }
    cross3(e[0], e[1], n1); if (!project6(n1, p, q)) return 0;
    cross3(f[0], f[1], m1); if (!project6(m1, p, q)) return 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { cross3(e[i], f[j], ax); if (!project6(ax, p, q)) return 0; }
    for (int i = 0; i < 3; ++i) { cross3(e[i], n1, ax); if (!project6(ax, p, q)) return 0; }
    for (int i = 0; i < 3; ++i) { cross3(f[i], m1, ax); if (!project6(ax, p, q)) return 0; }
    return 1;
}

/* conservative oriented-box vs node-box overlap: 3 world axes + the 3 box axes (false positives only) */
typedef struct { double c[3], ax[3][3] /* half-axis vectors */, ext[3] /* world half extents */; } OBox;

static int obox_hits_node(const OBox *b, const Node *n) {
    double e[3], d[3];
    for (int k = 0; k < 3; ++k) { e[k] = 0.5 * (n->hi[k] - n->lo[k]); d[k] = 0.5 * (n->hi[k] + n->lo[k]) - b->c[k]; }
    for (int k = 0; k < 3; ++k) if (fabs(d[k]) > e[k] + b->ext[k]) return 0;
    for (int j = 0; j < 3; ++j) {
        const double *a = b->ax[j];
        double len2 = dot3(a, a);
        double proj = fabs(dot3(d, a));
        double r_box = len2;                                       /* |a_j . a_j|; other axes are orthogonal up to 1e-7 */
        double r_node = e[0] * fabs(a[0]) + e[1] * fabs(a[1]) + e[2] * fabs(a[2]);
        if (proj > (r_box * 1.000001 + 1e-12) + r_node) return 0;
    }
    return 1;
}

/* does any of the part's posed triangles Q[0..nq) intersect a scene triangle? (part box walks the tree) */
static int part_hits_scene(const CoMesh *s, const OBox *b, const double *Q, int nq) {
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Node *n = s->nodes + stack[--sp];
        if (!obox_hits_node(b, n)) continue;
        if (n->count > 0) {
            for (int i = n->start; i < n->start + n->count; ++i) {
                int f = s->order[i];
                double P[9];
                for (int v = 0; v < 3; ++v) memcpy(P + 3 * v, s->V + 3 * s->F[3 * f + v], 24);
                for (int t = 0; t < nq; ++t) if (tri_tri(P, Q + 9 * t)) return 1;
            }
        } else { stack[sp++] = n->left; stack[sp++] = n->right; }
    }
    return 0;
}

/* parts: n_parts+1 triangle boundaries; each part's TCP-frame AABB becomes an oriented box under pose*diag(s,1,1) */
static int grasp_collides(const CoMesh *scene, const float *g, const double *gtris, const int *parts, int n_parts,
                          float ow, float tol, int use_width) {
    float sf = use_width ? (g[13] + tol) / ow : 1.0f;       /* float32, numpy >= 2 promotion of sampling.py:394 */
    double s = (double)sf, M[3][3], T[3] = {g[0], g[1], g[2]};
    for (int i = 0; i < 3; ++i) { M[i][0] = (double)g[3 + 3 * i] * s; M[i][1] = g[4 + 3 * i]; M[i][2] = g[5 + 3 * i]; }
    for (int p = 0; p < n_parts; ++p) {
        int t0 = parts[p], t1 = parts[p + 1], nq = t1 - t0;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        double *Q = (double *)malloc(sizeof(double) * 9 * nq);
        for (int t = t0; t < t1; ++t)
            for (int v = 0; v < 3; ++v) {
                const double *pt = gtris + 9 * t + 3 * v;
                for (int k = 0; k < 3; ++k) { if (pt[k] < lo[k]) lo[k] = pt[k]; if (pt[k] > hi[k]) hi[k] = pt[k]; }
                for (int i = 0; i < 3; ++i) Q[9 * (t - t0) + 3 * v + i] = ((M[i][0] * pt[0] + M[i][1] * pt[1]) + M[i][2] * pt[2]) + T[i];
            }
        OBox b;
        double cl[3], hl[3];
        for (int k = 0; k < 3; ++k) { cl[k] = 0.5 * (lo[k] + hi[k]); hl[k] = 0.5 * (hi[k] - lo[k]) * 1.000001 + 1e-9; }
        for (int i = 0; i < 3; ++i) {
            b.c[i] = ((M[i][0] * cl[0] + M[i][1] * cl[1]) + M[i][2] * cl[2]) + T[i];
            for (int j = 0; j < 3; ++j) b.ax[j][i] = M[i][j] * hl[j];
        }
        for (int i = 0; i < 3; ++i) b.ext[i] = (fabs(b.ax[0][i]) + fabs(b.ax[1][i]) + fabs(b.ax[2][i])) * 1.000001 + 1e-9;
        int hit = part_hits_scene(scene, &b, Q, nq);
        free(Q);
        if (hit) return 1;
    }
    return 0;
}

void co_collide(const CoMesh *scene, const float *gs, int64_t n, const double *gtris, const int *parts, int n_parts,
                double ow, double tol, int use_width, uint8_t *colliding) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; ++i)
        colliding[i] = (uint8_t)grasp_collides(scene, gs + 14 * i, gtris, parts, n_parts, (float)ow, (float)tol, use_width);
}

/* ---- the whole per-reference-point pipeline (what bench.py --impl reference times) --------------------------
 * per reference point: cast n_rays rays, filter, take the FIRST valid hit in canonical order (max_targets = 1),
 * expand n_o orientations, collision-check each posed gripper against `scene` (NULL = skip).
 * rows [n_ref*n_o,14], contacts [n_ref*n_o,6], colliding [n_ref*n_o], pair_count [n_ref]; rows of reference points
 * without a valid target are left zero.  returns the number of contact pairs found. */
int64_t co_pipeline(const CoMesh *m, const CoMesh *scene, const double *ref_pts, const double *dirs, const double *frame_vecs,
                    int64_t n_ref, int n_rays, const CoParams *P, int n_o, const double *cs, const double *sn,
                    const double *gtris, const int *parts, int n_parts, double coll_tol, int use_width,
                    float *rows, double *contacts, uint8_t *colliding, int32_t *pair_count) {
    int64_t total = 0;
    const int cap = P->max_hits_per_ray;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int64_t r = 0; r < n_ref; ++r) {
        CoHit *buf = (CoHit *)malloc(sizeof(CoHit) * cap);
        const double *p_r = ref_pts + 3 * r;
        int found = 0;
        double q[3];
        for (int ray = 0; ray < n_rays; ++ray) {
            int c = cast_one(m, p_r, dirs + 3 * (r * n_rays + ray), P, buf, cap);
            for (int h = 0; h < c && !found; ++h)
                if (buf[h].flags == 0) { found = 1; memcpy(q, buf[h].loc, 24); }
            /* the reference evaluates all rays before choosing; keep casting so the work is the same */
        }
        pair_count[r] = found;
        if (found) {
            expand_one(p_r, q, frame_vecs + 3 * r, n_o, cs, sn, rows + 14 * n_o * r, contacts ? contacts + 6 * n_o * r : NULL);
            if (scene)
                for (int o = 0; o < n_o; ++o)
                    colliding[n_o * r + o] = (uint8_t)grasp_collides(scene, rows + 14 * (n_o * r + o), gtris, parts, n_parts,
                                                                     (float)P->opening_width, (float)coll_tol, use_width);
            total += 1;
        }
        free(buf);
    }
    return total;
}
