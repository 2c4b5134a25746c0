/* TEST INFRASTRUCTURE ONLY — never part of the product path.
 *
 * pt_oracle.c: a plain-C restatement of the reference's per-pixel path-tracing hot path
 * (RenderKernel::render -> ray_trace_pixel and everything under it). Citations are file:line in /root/reference.
 *
 * PARITY PIN: tests/test_oracle_pin.py checks this file against
 *   (1) the reference's own golden vectors (include/bvh_tests.h: 572 hit rays + points, 222 miss rays), and
 *   (2) the compiled reference itself (oracle/_ref/libref_oracle.so): primary-ray prim/t maps and full render()
 *       framebuffers must be BIT-IDENTICAL on the bundled scenes (both are built -O2 -ffp-contract=off, no -march,
 *       against the same glibc libm), and fixtures generated from it are committed under tests/golden/.
 *
 * Every float expression below keeps the reference's operand order and its float/double promotions
 * (the "dbl" notes), because that is what makes (2) hold. The only deliberate difference is the acceleration
 * structure: the reference walks a pointer octree with 7-slab volumes (include/bvh.h:143-209); this file walks a
 * median-split AABB BVH with conservative box tests. Both return the exact closest hit over all triangles
 * (the octree equals brute force on every probed ray), so the result is the same; exact-t ties go to
 * the lowest triangle index, as in the reference's brute-force intersect_scene (render_kernel.cpp:453-483).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_1_PI
#define M_1_PI 0.31830988618379067154
#endif

typedef struct { float x, y, z; } v3;
typedef struct { float r, g, b; } col;   /* the reference Color carries a junk alpha (color.h:14,144-147): dropped */

/* ---- include/vec.h:75-189 ---------------------------------------------------------------------------------- */
static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline v3 vscale(float k, v3 a) { return V(k * a.x, k * a.y, k * a.z); }          /* vec.h:141-149 */
static inline float vdot(v3 u, v3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }         /* vec.h:186-189 */
static inline v3 vcross(v3 u, v3 v)                                                        /* vec.h:178-184 */
{
    return V((u.y * v.z) - (u.z * v.y), (u.z * v.x) - (u.x * v.z), (u.x * v.y) - (u.y * v.x));
}
static inline float vlength(v3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }     /* vec.h:157-165 */
static inline v3 vnormalize(v3 v) { float kk = 1.0f / vlength(v); return vscale(kk, v); }  /* vec.h:172-176 */
static inline v3 vdiv(v3 v, float k) { float kk = 1.0f / k; return vscale(kk, v); }        /* vec.h:151-155 */
/* std::max / std::min exactly as libstdc++ defines them (NaN behaviour included) */
static inline float smax(float a, float b) { return (a < b) ? b : a; }
static inline float smin(float a, float b) { return (b < a) ? b : a; }

/* ---- include/color.h ------------------------------------------------------------------------------------------ */
static inline col CO(float r, float g, float b) { col c = { r, g, b }; return c; }
static inline col cadd(col a, col b) { return CO(a.r + b.r, a.g + b.g, a.b + b.b); }       /* color.h:128-131 */
static inline col cmul(col a, col b) { return CO(a.r * b.r, a.g * b.g, a.b * b.b); }       /* color.h:144-147 */
static inline col cscale(col c, float k) { return CO(c.r * k, c.g * k, c.b * k); }         /* color.h:149-157 */
static inline col cdivf(col c, float k) { float kk = 1.0f / k; return cscale(c, kk); }     /* color.h:169-173 */
static inline int cblack(col c) { return c.r == 0.0f && c.g == 0.0f && c.b == 0.0f; }      /* color.h:23-26 */

/* ---- scene ------------------------------------------------------------------------------------------------------ */
typedef struct { v3 a, b, c; } tri_t;                                      /* include/triangle.h:8-68 */
typedef struct { col emission, diffuse; float metalness, roughness; } mat_t; /* include/simple_material.h:6-13 */
typedef struct { v3 center; float radius; int prim; } sph_t;               /* include/sphere.h:7-59 */
typedef struct { v3 p, n; float t, u, v; int prim; } hit_t;                /* include/hit_info.h:6-15 */

typedef struct { float lo[3], hi[3]; int left, right, first, count; } bnode_t;

typedef struct
{
    tri_t* tris; int n_tri;
    int* mat_idx; int n_mat_idx;
    mat_t* mats; int n_mat;
    int* emissive; int n_emissive;
    sph_t* spheres; int n_sph;
    float* env; int env_w, env_h;       /* rgba rows */
    float* cdf;
    bnode_t* nodes; int n_nodes; int* order;
} scene_t;

typedef struct { float m[4][4]; float fov_dist; int w, h; } camera_t;

/* ---- include/xorshift.h:10-31 ------------------------------------------------------------------------------------ */
static inline uint32_t xs_next(uint32_t* s)
{
    uint32_t x = *s;
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    return *s = x;
}
static inline float xs_float(uint32_t* s)
{
    /* (float)UINT_MAX rounds to 2^32; the cap is 1.0f - 1.0e-6f (xorshift.h:17) */
    return smin((float)xs_next(s) / (float)4294967295u, 1.0f - 1.0e-6f);
}

/* ---- source/mat.cpp:94-111 Transform::operator()(Point) ---------------------------------------------------------- */
static v3 xform_point(const float m[4][4], v3 p)
{
    float xt = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
    float yt = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
    float zt = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
    float wt = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
    float w = 1.f / wt;
    if (wt == 1.f) return V(xt, yt, zt);
    return V(xt * w, yt * w, zt * w);
}

/* ---- source/render_kernel.cpp:56-73 get_camera_ray ---------------------------------------------------------------- */
static void camera_ray(const camera_t* c, float x, float y, v3* o, v3* d)
{
    float x_ndc = x / c->w * 2 - 1;
    x_ndc *= (float)c->w / c->h;
    float y_ndc = y / c->h * 2 - 1;
    v3 origin = xform_point(c->m, V(0.0f, 0.0f, 0.0f));
    v3 pw = xform_point(c->m, V(x_ndc, y_ndc, c->fov_dist));
    *o = origin;
    *d = vnormalize(vsub(pw, origin));
}

/* ---- include/triangle.h:16-60 Moller-Trumbore --------------------------------------------------------------------- */
static inline int tri_intersect(const tri_t* tr, v3 o, v3 d, hit_t* h)
{
    const float EPSILON = 0.0000001f;
    v3 edge1 = vsub(tr->b, tr->a);
    v3 edge2 = vsub(tr->c, tr->a);
    v3 hh = vcross(d, edge2);
    float a = vdot(edge1, hh);
    if (a > -EPSILON && a < EPSILON) return 0;
    float f = 1.0f / a;
    v3 s = vsub(o, tr->a);
    float u = f * vdot(s, hh);
    if (u < 0.0f || u > 1.0f) return 0;
    v3 q = vcross(s, edge1);
    float v = f * vdot(d, q);
    if (v < 0.0f || u + v > 1.0f) return 0;
    float t = f * vdot(edge2, q);
    if (t > EPSILON)
    {
        h->p = vadd(o, vscale(t, d));
        h->n = vnormalize(vcross(edge1, edge2));
        h->t = t; h->u = u; h->v = v;
        return 1;
    }
    return 0;
}

/* ---- include/sphere.h:11-53 ----------------------------------------------------------------------------------------- */
static inline int sph_intersect(const sph_t* s, v3 o, v3 d, hit_t* h)
{
    v3 L = vsub(o, s->center);
    float b = 2.0f * vdot(d, L);
    float c = vdot(L, L) - s->radius * s->radius;
    float delta = b * b - 4.0f * 1.0f * c;
    if (delta < 0.0f) return 0;
    if (delta == 0.0f) h->t = -b / 2.0f;
    else
    {
        float sq = sqrtf(delta);
        float t1 = (-b - sq) / 2.0f;
        float t2 = (-b + sq) / 2.0f;
        if (t1 < t2) { h->t = t1; if (h->t < 0.0f) h->t = t2; }
    }
    if (h->t < 0.0f) return 0;
    h->p = vadd(o, vscale(h->t, d));
    h->n = vnormalize(vsub(h->p, s->center));
    h->prim = s->prim;
    return 1;
}

static inline hit_t hit_default(void)            /* include/hit_info.h:6-15 */
{
    hit_t h; h.p = V(0, 0, 0); h.n = V(0, 0, 0); h.t = -1.0f; h.u = -1; h.v = -1; h.prim = -1; return h;
}

/* ---- closest hit over the triangles: the port's own AABB BVH (see header) ------------------------------------------ */
static inline void consider(const scene_t* s, int ti, v3 o, v3 d, hit_t* best)
{
    hit_t h;
    if (tri_intersect(&s->tris[ti], o, d, &h))
        /* strict "<" as bvh.h:158 / render_kernel.cpp:464; equal t resolves to the lower index (brute-force order) */
        if (best->t == -1.0f || h.t < best->t || (h.t == best->t && ti < best->prim)) { *best = h; best->prim = ti; }
}

static void tris_brute(const scene_t* s, v3 o, v3 d, hit_t* best)
{
    for (int i = 0; i < s->n_tri; i++) consider(s, i, o, d, best);
}

static inline int box_hit(const bnode_t* n, v3 o, v3 inv, float tbest, float* tn_out)
{
    float tn = -INFINITY, tf = INFINITY;
    const float oo[3] = { o.x, o.y, o.z }, ii[3] = { inv.x, inv.y, inv.z };
    for (int a = 0; a < 3; a++)
    {
        float t0 = (n->lo[a] - oo[a]) * ii[a];
        float t1 = (n->hi[a] - oo[a]) * ii[a];
        tn = fmaxf(tn, fminf(t0, t1));      /* fminf/fmaxf drop NaNs (0 * inf): keeps the test conservative */
        tf = fminf(tf, fmaxf(t0, t1));
    }
    *tn_out = tn;
    if (tf < 0.0f) return 0;                       /* box behind the origin: a hit needs t > 1e-7 (triangle.h:44) */
    tf *= 1.0000005f;                              /* Ize-style widening on top of the build-time padding */
    if (tn > tf) return 0;
    if (tbest != -1.0f && tn > tbest) return 0;    /* strict: equal-t candidates are still tested (tie rule) */
    return 1;
}

static void tris_bvh(const scene_t* s, v3 o, v3 d, hit_t* best)
{
    if (s->n_tri == 0) return;
    v3 inv = V(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp)
    {
        const bnode_t* n = &s->nodes[stack[--sp]];
        float tn;
        if (!box_hit(n, o, inv, best->t, &tn)) continue;
        if (n->count)
        {
            for (int i = 0; i < n->count; i++) consider(s, s->order[n->first + i], o, d, best);
            continue;
        }
        float tl, trr;
        int hl = box_hit(&s->nodes[n->left], o, inv, best->t, &tl);
        int hr = box_hit(&s->nodes[n->right], o, inv, best->t, &trr);
        if (hl && hr)
        {
            if (tl <= trr) { stack[sp++] = n->right; stack[sp++] = n->left; }
            else { stack[sp++] = n->left; stack[sp++] = n->right; }
        }
        else if (hl) stack[sp++] = n->left;
        else if (hr) stack[sp++] = n->right;
    }
}

/* render_kernel.cpp:485-502 intersect_scene_bvh / :453-483 intersect_scene; returns closest.t > 0 */
static int intersect_scene(const scene_t* s, v3 o, v3 d, hit_t* closest, int brute)
{
    if (brute) tris_brute(s, o, d, closest); else tris_bvh(s, o, d, closest);
    for (int i = 0; i < s->n_sph; i++)
    {
        hit_t h = hit_default();
        if (sph_intersect(&s->spheres[i], o, d, &h))
            if (h.t < closest->t || closest->t == -1.0f) *closest = h;
    }
    return closest->t > 0.0f;
}

/* ---- BVH build (port-only; median split over centroid bounds) -------------------------------------------------------- */
typedef struct { float c[3]; float lo[3], hi[3]; } prim_t;
static prim_t* g_prims;   /* build is single-threaded */

static int build_rec(scene_t* s, int* order, int first, int count, int* n_nodes, int depth)
{
    int me = (*n_nodes)++;
    bnode_t* n = &s->nodes[me];
    float clo[3] = { INFINITY, INFINITY, INFINITY }, chi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int a = 0; a < 3; a++) { n->lo[a] = INFINITY; n->hi[a] = -INFINITY; }
    for (int i = first; i < first + count; i++)
    {
        const prim_t* p = &g_prims[order[i]];
        for (int a = 0; a < 3; a++)
        {
            n->lo[a] = fminf(n->lo[a], p->lo[a]); n->hi[a] = fmaxf(n->hi[a], p->hi[a]);
            clo[a] = fminf(clo[a], p->c[a]); chi[a] = fmaxf(chi[a], p->c[a]);
        }
    }
    for (int a = 0; a < 3; a++)
    {   /* pad: the box test must never reject a box whose triangle the exact test would accept */
        float pad = 1e-5f * fmaxf(fmaxf(fabsf(n->lo[a]), fabsf(n->hi[a])), 1.0f);
        n->lo[a] -= pad; n->hi[a] += pad;
    }
    n->left = n->right = -1; n->first = first; n->count = 0;
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    if (count <= 4 || depth > 100 || !(chi[axis] > clo[axis])) { n->count = count; return me; }
    /* partition around the spatial middle; fall back to an even split when one side is empty */
    float mid = 0.5f * (clo[axis] + chi[axis]);
    int i = first, j = first + count - 1;
    while (i <= j)
    {
        if (g_prims[order[i]].c[axis] < mid) i++;
        else { int tmp = order[i]; order[i] = order[j]; order[j] = tmp; j--; }
    }
    int nl = i - first;
    if (nl == 0 || nl == count) nl = count / 2;
    int l = build_rec(s, order, first, nl, n_nodes, depth + 1);
    int r = build_rec(s, order, first + nl, count - nl, n_nodes, depth + 1);
    n = &s->nodes[me];
    n->left = l; n->right = r;
    return me;
}

static void build_bvh(scene_t* s)
{
    if (s->n_tri == 0) { s->nodes = NULL; s->order = NULL; s->n_nodes = 0; return; }
    g_prims = (prim_t*)malloc(sizeof(prim_t) * (size_t)s->n_tri);
    s->order = (int*)malloc(sizeof(int) * (size_t)s->n_tri);
    s->nodes = (bnode_t*)malloc(sizeof(bnode_t) * (size_t)(2 * s->n_tri + 1));
    for (int i = 0; i < s->n_tri; i++)
    {
        const tri_t* t = &s->tris[i];
        const float vx[3][3] = { { t->a.x, t->a.y, t->a.z }, { t->b.x, t->b.y, t->b.z }, { t->c.x, t->c.y, t->c.z } };
        for (int a = 0; a < 3; a++)
        {
            g_prims[i].lo[a] = fminf(vx[0][a], fminf(vx[1][a], vx[2][a]));
            g_prims[i].hi[a] = fmaxf(vx[0][a], fmaxf(vx[1][a], vx[2][a]));
            g_prims[i].c[a] = 0.5f * (g_prims[i].lo[a] + g_prims[i].hi[a]);
        }
        s->order[i] = i;
    }
    int n_nodes = 0;
    build_rec(s, s->order, 0, s->n_tri, &n_nodes, 0);
    s->n_nodes = n_nodes;
    free(g_prims); g_prims = NULL;
}

/* ---- include/image.h:80-85 (dbl), :165-177 ------------------------------------------------------------------------- */
static inline int env_offset(const scene_t* s, int x, int y)
{
    int px = x; if (px < 0) px = 0; if (px > s->env_w - 1) px = s->env_w - 1;
    int py = y; if (py < 0) py = 0; if (py > s->env_h - 1) py = s->env_h - 1;
    return py * s->env_w + px;
}
static inline col env_texel(const scene_t* s, int idx) { const float* p = s->env + 4 * (size_t)idx; return CO(p[0], p[1], p[2]); }
static inline float env_luminance_of_pixel(const scene_t* s, int x, int y)
{
    col p = env_texel(s, env_offset(s, x, y));
    return (float)(0.3086 * p.r + 0.6094 * p.g + 0.0820 * p.b);   /* dbl: double constants, image.h:84 */
}

/* ---- source/utils.cpp:126-142 compute_env_map_cdf ------------------------------------------------------------------- */
static void compute_env_cdf(scene_t* s)
{
    int n = s->env_w * s->env_h;
    s->cdf[0] = 0.0f;
    for (int y = 0; y < s->env_h; y++)
        for (int x = 0; x < s->env_w; x++)
        {
            int index = y * s->env_w + x;
            int prev = index - 1 > 0 ? index - 1 : 0;
            s->cdf[index] = s->cdf[prev] + env_luminance_of_pixel(s, x, y);
        }
    (void)n;
}

/* ---- BRDF: render_kernel.cpp:5-22, :218-301, :392-451, :513-518 ----------------------------------------------------- */
static void branchless_onb(v3 n, v3* b1, v3* b2)                                   /* :5-12 */
{
    float sign = copysignf(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float b = n.x * n.y * a;
    *b1 = V(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    *b2 = V(b, sign + n.y * n.y * a, -n.y);
}
static v3 rotate_around_normal(v3 normal, v3 local)                                /* :14-22 */
{
    v3 t, bt;
    branchless_onb(normal, &t, &bt);
    return vadd(vadd(vscale(local.x, t), vscale(local.y, bt)), vscale(local.z, normal));
}
static col fresnel_schlick(col F0, float NoV)                                      /* :218-221 */
{
    float p = powf((1.0f - NoV), 5.0f);
    return CO(F0.r + (1.0f + -F0.r) * p, F0.g + (1.0f + -F0.g) * p, F0.b + (1.0f + -F0.b) * p);
}
static float ggx_d(float alpha, float NoH)                                         /* :223-233 */
{
    NoH = smin(NoH, 0.999999f);
    float alpha2 = alpha * alpha;
    float NoH2 = NoH * NoH;
    float b = (NoH2 * (alpha2 - 1.0f) + 1.0f);
    return (float)(alpha2 * M_1_PI / (b * b));                                    /* dbl: :232 */
}
static float g1_schlick(float k, float dp) { return dp / (dp * (1.0f - k) + k); }  /* :235-238 */
static float ggx_smith(float rough2, float NoV, float NoL)                         /* :240-245 */
{
    float k = rough2 / 2.0f;
    return g1_schlick(k, NoL) * g1_schlick(k, NoV);
}
static float power_heuristic(float a, float b) { float a2 = a * a; return a2 / (a2 + b * b); }   /* :513-518 */

static float ct_pdf(const mat_t* m, v3 view, v3 to_light, v3 n)                    /* :247-258 */
{
    v3 h = vnormalize(vadd(view, to_light));
    float alpha = m->roughness * m->roughness;
    float VoH = smax(0.0f, vdot(view, h));
    float NoH = smax(0.0f, vdot(n, h));
    float D = ggx_d(alpha, NoH);
    return D * NoH / (4.0f * VoH);
}

/* the block shared by cook_torrance_brdf (:272-298) and the importance sampler (:424-448) */
static col ct_eval_terms(const mat_t* m, float NoV, float NoL, float NoH, float VoH, float* pdf_out)
{
    float metalness = m->metalness;
    float alpha = m->roughness * m->roughness;
    col base = m->diffuse;
    float f04 = 0.04f * (1.0f - metalness);
    col F0 = CO(f04 + base.r * metalness, f04 + base.g * metalness, f04 + base.b * metalness);   /* :284 */
    col F = fresnel_schlick(F0, VoH);
    float D = ggx_d(alpha, NoH);
    float G = ggx_smith(alpha, NoV, NoL);
    float kd0 = 1.0f - metalness;
    col kD = CO(kd0 * (1.0f + -F.r), kd0 * (1.0f + -F.g), kd0 * (1.0f + -F.b));                  /* :291-292 */
    col diffuse_part = cdivf(cmul(kD, base), (float)M_PI);                                       /* :294 */
    col specular_part = cdivf(cscale(cscale(F, D), G), 4.0f * NoV * NoL);                        /* :295 */
    if (pdf_out) *pdf_out = D * NoH / (4.0f * VoH);                                              /* :445 */
    return cadd(diffuse_part, specular_part);
}

static col ct_brdf(const mat_t* m, v3 to_light, v3 view, v3 n)                     /* :260-301 */
{
    v3 h = vnormalize(vadd(view, to_light));
    float NoV = smax(0.0f, vdot(n, view));
    float NoL = smax(0.0f, vdot(n, to_light));
    float NoH = smax(0.0f, vdot(n, h));
    float VoH = smax(0.0f, vdot(h, view));
    if (NoV > 0.0f && NoL > 0.0f && NoH > 0.0f)
        return ct_eval_terms(m, NoV, NoL, NoH, VoH, NULL);
    return CO(0.0f, 0.0f, 0.0f);
}

static col ct_importance_sample(const mat_t* m, v3 view, v3 n, v3* out_dir, float* pdf, uint32_t* rng)   /* :392-451 */
{
    *pdf = 0.0f;
    float alpha = m->roughness * m->roughness;
    float rand1 = xs_float(rng);
    float rand2 = xs_float(rng);
    float phi = 2.0f * (float)M_PI * rand1;
    float theta = acosf((1.0f - rand2) / (rand2 * (alpha * alpha - 1.0f) + 1.0f));     /* :404 (no sqrt) */
    float sin_theta = sinf(theta);
    v3 local = V(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cosf(theta));
    v3 mn = rotate_around_normal(n, local);
    if (vdot(mn, n) < 0.0f) return CO(0.0f, 0.0f, 0.0f);                               /* out_dir left untouched */
    v3 to_light = vnormalize(vsub(vscale(2.0f * vdot(mn, view), mn), view));
    *out_dir = to_light;
    float NoV = smax(0.0f, vdot(n, view));
    float NoL = smax(0.0f, vdot(n, to_light));
    float NoH = smax(0.0f, vdot(n, mn));
    float VoH = smax(0.0f, vdot(mn, view));
    if (NoV > 0.0f && NoL > 0.0f && NoH > 0.0f)
        return ct_eval_terms(m, NoV, NoL, NoH, VoH, pdf);
    return CO(0.0f, 0.0f, 0.0f);
}

/* ---- environment map: render_kernel.cpp:520-631 ---------------------------------------------------------------------- */
static col env_from_direction(const scene_t* s, v3 d)                              /* :520-530 */
{
    float u = 0.5f + atan2f(d.z, d.x) / (2.0f * (float)M_PI);
    float v = 0.5f + asinf(d.y) / (float)M_PI;
    int x = (int)(u * s->env_w); if (x > s->env_w - 1) x = s->env_w - 1; if (x < 0) x = 0;
    int y = (int)(v * s->env_h); if (y > s->env_h - 1) y = s->env_h - 1; if (y < 0) y = 0;
    return env_texel(s, y * s->env_w + x);
}

static void env_cdf_search(const scene_t* s, float value, int* xo, int* yo)        /* :532-567 */
{
    int lower = 0, upper = s->env_h - 1;
    int x_index = s->env_w - 1;
    while (lower < upper)
    {
        int y_index = (lower + upper) / 2;
        if (value < s->cdf[y_index * s->env_w + x_index]) upper = y_index; else lower = y_index + 1;
    }
    int y = lower; if (y > s->env_h) y = s->env_h; if (y < 0) y = 0;
    lower = 0; upper = s->env_w - 1;
    while (lower < upper)
    {
        int xi = (lower + upper) / 2;
        if (value < s->cdf[y * s->env_w + xi]) upper = xi; else lower = xi + 1;
    }
    int x = lower; if (x > s->env_w) x = s->env_w; if (x < 0) x = 0;
    *xo = x; *yo = y;
}

typedef struct { long long rays; int brute; } ctx_t;

static inline int trace(const scene_t* s, ctx_t* c, v3 o, v3 d, hit_t* h)          /* INTERSECT_SCENE :504-511 */
{
    c->rays++;
    *h = hit_default();
    return intersect_scene(s, o, d, h, c->brute);
}

static col sample_environment_map(const scene_t* s, ctx_t* c, v3 ray_d, const hit_t* hit, const mat_t* m, uint32_t* rng)  /* :569-631 */
{
    float total = s->cdf[s->env_w * s->env_h - 1];
    int x, y;
    env_cdf_search(s, xs_float(rng) * total, &x, &y);
    float u = (float)x / s->env_w;
    float v = (float)y / s->env_h;
    float phi = (float)(u * 2.0f * M_PI);                                           /* dbl: :578 */
    float theta = (float)(v * M_PI);                                                /* dbl: :579 */
    col env_sample = CO(0.0f, 0.0f, 0.0f);
    float sin_theta = sinf(theta);
    float cos_theta = cosf(theta);
    /* unqualified cos/sin on a float (:586): <math.h> is the libstdc++ wrapper (camera.h:6-7), so the float overloads win */
    v3 dir = V(-sin_theta * cosf(phi), -cos_theta, -sin_theta * sinf(phi));
    v3 view = vneg(ray_d);
    float cosine_term = vdot(hit->n, dir);
    if (cosine_term > 0.0f)
    {
        hit_t trash;
        if (!trace(s, c, vadd(hit->p, vscale(1.0e-4f, hit->n)), dir, &trash))
        {
            float env_pdf = env_luminance_of_pixel(s, x, y) / total;
            env_pdf = (float)((env_pdf * s->env_w * s->env_h) / (2.0f * M_PI * M_PI * sin_theta));   /* dbl: :595 */
            col radiance = env_texel(s, env_offset(s, x, y));
            col brdf = ct_brdf(m, dir, view, hit->n);
            float brdf_pdf = ct_pdf(m, view, dir, hit->n);
            float w = power_heuristic(env_pdf, brdf_pdf);
            env_sample = cdivf(cmul(cscale(cscale(brdf, cosine_term), w), radiance), env_pdf);       /* :602 */
        }
    }
    float bpdf; v3 bdir = V(0, 0, 0);
    col bsample_brdf = ct_importance_sample(m, view, hit->n, &bdir, &bpdf, rng);
    cosine_term = smax(vdot(hit->n, bdir), 0.0f);
    col brdf_sample = CO(0.0f, 0.0f, 0.0f);
    if (bpdf != 0.0f && cosine_term > 0.0f)
    {
        hit_t trash;
        if (!trace(s, c, vadd(hit->p, vscale(1.0e-5f, hit->n)), bdir, &trash))
        {
            col sky = env_from_direction(s, bdir);
            float theta_b = acosf(bdir.z);                                          /* :618 (z, not y: kept) */
            float sin_b = sinf(theta_b);
            float env_pdf = (0.3086f * sky.r + 0.6094f * sky.g + 0.0820f * sky.b) / total;          /* color.h:89-92 */
            env_pdf *= s->env_w * s->env_h;                                         /* int product, :622 */
            env_pdf = (float)(env_pdf / (2.0f * M_PI * M_PI * sin_b));              /* dbl: :623 */
            float w = power_heuristic(bpdf, env_pdf);
            brdf_sample = cdivf(cmul(cscale(cscale(sky, w), cosine_term), bsample_brdf), bpdf);      /* :626 */
        }
    }
    return cadd(brdf_sample, env_sample);
}

/* ---- emissive triangles: render_kernel.cpp:633-759 -------------------------------------------------------------------- */
static col sample_light_sources(const scene_t* s, ctx_t* c, v3 ray_d, const hit_t* hit, const mat_t* m, uint32_t* rng)
{
    col light_mis = CO(0.0f, 0.0f, 0.0f);
    v3 view = vneg(ray_d);
    if (s->n_emissive > 0)
    {
        /* sample_random_point_on_lights :715-742 */
        int pick = (int)(xs_float(rng) * (float)(size_t)s->n_emissive);
        int em_tri = s->emissive[pick];
        const tri_t* lt = &s->tris[em_tri];
        float rand_1 = xs_float(rng);
        float rand_2 = xs_float(rng);
        float sqrt_r1 = sqrtf(rand_1);
        float u = 1.0f - sqrt_r1;
        float v = (1.0f - rand_2) * sqrt_r1;
        v3 AB = vsub(lt->b, lt->a);
        v3 AC = vsub(lt->c, lt->a);
        v3 lp = vadd(vadd(lt->a, vscale(u, AB)), vscale(v, AC));
        v3 nrm = vcross(AB, AC);
        float len_n = vlength(nrm);
        v3 light_n = vdiv(nrm, len_n);
        float area = len_n * 0.5f;
        float light_pdf = 1.0f / ((float)(size_t)s->n_emissive * area);

        v3 so = vadd(hit->p, vscale(1.0e-4f, hit->n));                              /* :642 */
        v3 sd = vsub(lp, so);
        float dist = vlength(sd);
        v3 sdn = vnormalize(sd);
        float dot_light = smax(vdot(light_n, vneg(sdn)), 0.0f);
        if (dot_light > 0.0f)
        {
            hit_t sh;                                                               /* evaluate_shadow_ray :744-759 */
            int found = trace(s, c, so, sdn, &sh);
            int in_shadow = found && (sh.t + 1.0e-4f < dist);
            if (!in_shadow)
            {
                const mat_t* em = &s->mats[s->mat_idx[em_tri]];
                light_pdf *= dist * dist;
                light_pdf /= dot_light;
                col brdf = ct_brdf(m, sdn, view, hit->n);
                float bp = ct_pdf(m, view, sdn, hit->n);
                if (bp != 0.0f)
                {
                    float w = power_heuristic(light_pdf, bp);
                    float cosine_term = vdot(hit->n, sdn);
                    light_mis = cdivf(cscale(cmul(cscale(em->emission, cosine_term), brdf), w), light_pdf);   /* :671 */
                }
            }
        }
    }
    col brdf_mis = CO(0.0f, 0.0f, 0.0f);
    v3 bdir = V(0, 0, 0); float dpdf;
    col brdf = ct_importance_sample(m, view, hit->n, &bdir, &dpdf, rng);            /* :681 */
    if (!cblack(brdf))
    {
        hit_t nh;
        if (trace(s, c, vadd(hit->p, vscale(1.0e-5f, hit->n)), bdir, &nh))          /* :684-686 */
        {
            float cos_angle = smax(vdot(nh.n, vneg(bdir)), 0.0f);
            if (cos_angle > 0.0f)
            {
                const mat_t* hm = &s->mats[s->mat_idx[nh.prim]];
                col emission = hm->emission;
                if (emission.r > 0 || emission.g > 0 || emission.b > 0)
                {
                    float d2 = nh.t * nh.t;
                    const tri_t* ht = &s->tris[nh.prim];
                    float light_area = vlength(vcross(vsub(ht->b, ht->a), vsub(ht->c, ht->a))) / 2;   /* triangle.cpp:8-11 */
                    float light_pdf = d2 / (light_area * cos_angle);                 /* :702 (no 1/N: kept) */
                    float w = power_heuristic(dpdf, light_pdf);
                    float cosine_term = vdot(hit->n, bdir);
                    brdf_mis = cdivf(cscale(cmul(cscale(brdf, cosine_term), emission), w), dpdf);     /* :706 */
                }
            }
        }
    }
    return cadd(light_mis, brdf_mis);
}

/* ---- source/render_kernel.cpp:75-181 ray_trace_pixel ------------------------------------------------------------------- */
static void trace_pixel(const scene_t* s, const camera_t* cam, int spp, int max_bounces, int x, int y, ctx_t* c, float* out4)
{
    uint32_t rng = (uint32_t)(31 + x * y * spp);                                    /* int arithmetic, wraps (-fwrapv) */
    for (int i = 0; i < 10; i++) xs_float(&rng);
    col final_color = CO(0.0f, 0.0f, 0.0f);
    for (int sample = 0; sample < spp; sample++)
    {
        float xj = (x + 0.5f) + xs_float(&rng) - 1.0f;
        float yj = (y + 0.5f) + xs_float(&rng) - 1.0f;
        v3 o, d;
        camera_ray(cam, xj, yj, &o, &d);
        col throughput = CO(1.0f, 1.0f, 1.0f);
        col sample_color = CO(0.0f, 0.0f, 0.0f);
        int state = 0;   /* 0 BOUNCE, 1 MISSED, 2 TERMINATED (ray.h:6-11) */
        for (int bounce = 0; bounce < max_bounces; bounce++)
        {
            if (state == 0)
            {
                hit_t hit;
                if (trace(s, c, o, d, &hit))
                {
                    const mat_t m = s->mats[s->mat_idx[hit.prim]];
                    col light = sample_light_sources(s, c, d, &hit, &m, &rng);
                    col env = sample_environment_map(s, c, d, &hit, &m, &rng);
                    float bpdf; v3 ndir = V(0, 0, 0);
                    col brdf = ct_importance_sample(&m, vneg(d), hit.n, &ndir, &bpdf, &rng);
                    if (bounce == 0) sample_color = cadd(sample_color, m.emission);
                    sample_color = cadd(sample_color, cmul(cadd(light, env), throughput));
                    if (cblack(brdf) || bpdf < 1.0e-8f || isinf(bpdf)) { state = 2; break; }
                    col f = cdivf(cscale(brdf, smax(0.0f, vdot(ndir, hit.n))), bpdf);                 /* :137 */
                    throughput = cmul(throughput, f);
                    o = vadd(hit.p, vscale(1.0e-4f, hit.n));
                    d = ndir;
                    state = 0;
                }
                else state = 1;
            }
            else if (state == 1)
            {
                if (bounce == 1)
                    sample_color = cadd(sample_color, cmul(env_from_direction(s, d), throughput));    /* :148-156 */
                break;
            }
            else break;
        }
        final_color = cadd(final_color, sample_color);
    }
    float n = (float)spp;
    final_color = CO(final_color.r / n, final_color.g / n, final_color.b / n);      /* operator/= :67-74: true division */
    /* framebuffer starts at Color::Black() (alpha 1): fb += final (alpha: operator+= leaves it, :41-48) then tone-map :171-180 */
    const float gamma = 2.2f, exposure = 1.5f;
    col hdr = CO(0.0f + final_color.r, 0.0f + final_color.g, 0.0f + final_color.b);
    col tm = CO(1.0f + -expf(-hdr.r * exposure), 1.0f + -expf(-hdr.g * exposure), 1.0f + -expf(-hdr.b * exposure));
    out4[0] = powf(tm.r, 1.0f / gamma);
    out4[1] = powf(tm.g, 1.0f / gamma);
    out4[2] = powf(tm.b, 1.0f / gamma);
    /* alpha: 1 (Black) -> exp() keeps col.a (-1*1.5 = -1.5), 1 + -(-1.5)... the reference ends at 2.5 */
    out4[3] = 2.5f;
}

/* ================================================ C ABI (same shape as oracle/ref_driver.cpp) ========================= */
int pto_max_threads(void) { return omp_get_max_threads(); }

void* pto_scene_from_arrays(const float* tri9, int n_tri, const int* mat_idx, const float* mats10, int n_mat,
                            const int* emissive, int n_emissive, const float* spheres4, const int* sphere_prim, int n_sph,
                            const int* extra_mat_idx, int n_extra)
{
    scene_t* s = (scene_t*)calloc(1, sizeof(scene_t));
    s->n_tri = n_tri;
    s->tris = (tri_t*)malloc(sizeof(tri_t) * (size_t)(n_tri > 0 ? n_tri : 1));
    for (int i = 0; i < n_tri; i++)
    {
        const float* p = tri9 + 9 * (size_t)i;
        s->tris[i].a = V(p[0], p[1], p[2]); s->tris[i].b = V(p[3], p[4], p[5]); s->tris[i].c = V(p[6], p[7], p[8]);
    }
    s->n_mat_idx = n_tri + n_extra;
    s->mat_idx = (int*)malloc(sizeof(int) * (size_t)(s->n_mat_idx > 0 ? s->n_mat_idx : 1));
    if (n_tri) memcpy(s->mat_idx, mat_idx, sizeof(int) * (size_t)n_tri);
    for (int i = 0; i < n_extra; i++) s->mat_idx[n_tri + i] = extra_mat_idx[i];
    s->n_mat = n_mat;
    s->mats = (mat_t*)malloc(sizeof(mat_t) * (size_t)(n_mat > 0 ? n_mat : 1));
    for (int i = 0; i < n_mat; i++)
    {
        const float* m = mats10 + 10 * (size_t)i;
        s->mats[i].emission = CO(m[0], m[1], m[2]); s->mats[i].diffuse = CO(m[4], m[5], m[6]);
        s->mats[i].metalness = m[8]; s->mats[i].roughness = m[9];
    }
    s->n_emissive = n_emissive;
    s->emissive = (int*)malloc(sizeof(int) * (size_t)(n_emissive > 0 ? n_emissive : 1));
    if (n_emissive) memcpy(s->emissive, emissive, sizeof(int) * (size_t)n_emissive);
    s->n_sph = n_sph;
    s->spheres = (sph_t*)malloc(sizeof(sph_t) * (size_t)(n_sph > 0 ? n_sph : 1));
    for (int i = 0; i < n_sph; i++)
    {
        s->spheres[i].center = V(spheres4[4 * i], spheres4[4 * i + 1], spheres4[4 * i + 2]);
        s->spheres[i].radius = spheres4[4 * i + 3];
        s->spheres[i].prim = sphere_prim[i];
    }
    build_bvh(s);
    return s;
}

void pto_scene_free(void* h)
{
    scene_t* s = (scene_t*)h;
    if (!s) return;
    free(s->tris); free(s->mat_idx); free(s->mats); free(s->emissive); free(s->spheres);
    free(s->env); free(s->cdf); free(s->nodes); free(s->order); free(s);
}

void pto_scene_counts(void* h, int* out7)
{
    scene_t* s = (scene_t*)h;
    out7[0] = s->n_tri; out7[1] = s->n_mat; out7[2] = s->n_emissive; out7[3] = s->n_mat_idx; out7[4] = s->n_sph;
    out7[5] = s->env_w; out7[6] = s->env_h;
}

void pto_set_env(void* h, const float* rgba, int w, int height)
{
    scene_t* s = (scene_t*)h;
    free(s->env); free(s->cdf);
    s->env_w = w; s->env_h = height;
    s->env = (float*)malloc(sizeof(float) * 4 * (size_t)w * height);
    memcpy(s->env, rgba, sizeof(float) * 4 * (size_t)w * height);
    s->cdf = (float*)malloc(sizeof(float) * (size_t)w * height);
    compute_env_cdf(s);
}

void pto_get_env_cdf(void* h, float* out)
{
    scene_t* s = (scene_t*)h;
    memcpy(out, s->cdf, sizeof(float) * (size_t)s->env_w * s->env_h);
}

static camera_t camera_from17(const float* cam17, int w, int h)
{
    camera_t c;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c.m[i][j] = cam17[4 * i + j];
    c.fov_dist = cam17[16]; c.w = w; c.h = h;
    return c;
}

void pto_camera_rays(const float* cam17, int w, int h, const float* xy, int n, float* rays6)
{
    camera_t c = camera_from17(cam17, w, h);
    for (int i = 0; i < n; i++)
    {
        v3 o, d;
        camera_ray(&c, xy[2 * i], xy[2 * i + 1], &o, &d);
        float* r = rays6 + 6 * (size_t)i;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
    }
}

/* mode 0/2: port BVH, triangles only; 1: brute force + spheres; 3: port BVH + spheres */
int pto_trace(void* h, const float* rays6, int n, int mode, int* prim, float* t, float* extra8, int nthreads)
{
    scene_t* s = (scene_t*)h;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++)
    {
        const float* r = rays6 + 6 * (size_t)i;
        v3 o = V(r[0], r[1], r[2]), d = V(r[3], r[4], r[5]);
        hit_t hit = hit_default();
        int found;
        if (mode == 0 || mode == 2) { tris_bvh(s, o, d, &hit); found = hit.t > 0.0f; }
        else found = intersect_scene(s, o, d, &hit, mode == 1);
        prim[i] = found ? hit.prim : -1;
        t[i] = found ? hit.t : -1.0f;
        if (extra8)
        {
            float* e = extra8 + 8 * (size_t)i;
            e[0] = hit.p.x; e[1] = hit.p.y; e[2] = hit.p.z; e[3] = hit.n.x; e[4] = hit.n.y; e[5] = hit.n.z; e[6] = hit.u; e[7] = hit.v;
        }
    }
    return 0;
}

double pto_primary(void* h, const float* cam17, int w, int height, int mode, int* prim, float* t, int nthreads)
{
    scene_t* s = (scene_t*)h;
    camera_t c = camera_from17(cam17, w, height);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < height; y++)
        for (int x = 0; x < w; x++)
        {
            v3 o, d;
            camera_ray(&c, (float)x, (float)y, &o, &d);
            hit_t hit = hit_default();
            int found;
            if (mode == 0 || mode == 2) { tris_bvh(s, o, d, &hit); found = hit.t > 0.0f; }
            else found = intersect_scene(s, o, d, &hit, mode == 1);
            prim[y * w + x] = found ? hit.prim : -1;
            t[y * w + x] = found ? hit.t : -1.0f;
        }
    return omp_get_wtime() - t0;
}

static double render_rect(scene_t* s, const float* cam17, int w, int height, int spp, int bounces,
                          int x0, int y0, int x1, int y1, float* out, int out_w, int ox, int oy, long long* rays, int nthreads)
{
    if (!s->cdf) return -1.0;
    camera_t c = camera_from17(cam17, w, height);
    if (nthreads > 0) omp_set_num_threads(nthreads);
    long long total = 0;
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic) reduction(+ : total)
    for (int y = y0; y < y1; y++)
    {
        ctx_t ctx = { 0, 0 };
        for (int x = x0; x < x1; x++)
            trace_pixel(s, &c, spp, bounces, x, y, &ctx, out + 4 * ((size_t)(y - oy) * out_w + (x - ox)));
        total += ctx.rays;
    }
    double sec = omp_get_wtime() - t0;
    if (rays) *rays = total;
    return sec;
}

double pto_render(void* h, const float* cam17, int w, int height, int spp, int bounces, float* out_rgba, int nthreads)
{
    return render_rect((scene_t*)h, cam17, w, height, spp, bounces, 0, 0, w, height, out_rgba, w, 0, 0, NULL, nthreads);
}

double pto_render_counted(void* h, const float* cam17, int w, int height, int spp, int bounces, float* out_rgba,
                          long long* rays, int nthreads)
{
    return render_rect((scene_t*)h, cam17, w, height, spp, bounces, 0, 0, w, height, out_rgba, w, 0, 0, rays, nthreads);
}

double pto_render_crop(void* h, const float* cam17, int w, int height, int spp, int bounces,
                       int x0, int y0, int x1, int y1, float* out_rgba_crop, int nthreads)
{
    return render_rect((scene_t*)h, cam17, w, height, spp, bounces, x0, y0, x1, y1, out_rgba_crop, x1 - x0, x0, y0, NULL, nthreads);
}

uint32_t pto_xorshift(uint32_t seed, int n_warmup, int n, float* out)
{
    uint32_t s = seed;
    for (int i = 0; i < n_warmup; i++) xs_float(&s);
    uint32_t after = s;
    for (int i = 0; i < n; i++) out[i] = xs_float(&s);
    return after;
}
