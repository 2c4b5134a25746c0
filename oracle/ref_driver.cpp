// TEST INFRASTRUCTURE ONLY — never part of the product path.
//
// C-ABI driver around the UNMODIFIED reference sources (compiled where they lie, from
// $REF_DIR = /root/reference, by oracle/Makefile into oracle/_ref/libref_oracle.so).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// Everything below calls the reference's own classes:
//   Utils::parse_obj            source/utils.cpp:16-98
//   Utils::compute_env_map_cdf  source/utils.cpp:126-142
//   BVH::BVH / BVH::intersect   source/bvh.cpp:19-65, include/bvh.h:127-209
//   BVH::flatten / FlattenedBVH::intersect   include/bvh.h:211-250, source/flattened_bvh.cpp:10-58
//   RenderKernel::{get_camera_ray, intersect_scene, intersect_scene_bvh, render}
//                               source/render_kernel.cpp:56-73, :453-502, :189-211
//   xorshift32_generator        include/xorshift.h:10-31
//   golden rays                 include/bvh_tests.h (572 hit rays + points, 222 miss rays)
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include <fcntl.h>
#include <omp.h>
#include <unistd.h>

#include "bvh.h"
#include "camera.h"
#include "image.h"
#include "render_kernel.h"
#include "utils.h"
#include "image_io.h"
#include "bvh_tests.h"

namespace {

struct RefScene
{
    std::vector<Triangle> triangles;
    std::vector<SimpleMaterial> materials;
    std::vector<int> emissive;
    std::vector<int> material_indices;
    std::vector<Sphere> spheres;
    BVH* bvh = nullptr;
    FlattenedBVH* flat = nullptr;
    Image env;
    std::vector<float> env_cdf;
    double bvh_build_seconds = 0.0;

    ~RefScene() { delete bvh; delete flat; }

    void build()
    {
        auto t0 = std::chrono::high_resolution_clock::now();
        bvh = new BVH(&triangles);
        auto t1 = std::chrono::high_resolution_clock::now();
        bvh_build_seconds = std::chrono::duration<double>(t1 - t0).count();
    }
};

Camera camera_from17(const float* cam17)
{
    Camera cam;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            cam.view_matrix.m[i][j] = cam17[i * 4 + j];
    cam.fov_dist = cam17[16];
    return cam;
}

void camera_to17(const Camera& cam, float* out17)
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            out17[i * 4 + j] = cam.view_matrix.m[i][j];
    out17[16] = cam.fov_dist;
}

// The reference prints a progress line per row from thread 0 (render_kernel.cpp:205-209).
struct StdoutSilencer
{
    int saved = -1;
    StdoutSilencer()
    {
        fflush(stdout);
        std::cout.flush();
        saved = dup(1);
        int devnull = open("/dev/null", O_WRONLY);
        if (devnull >= 0) { dup2(devnull, 1); close(devnull); }
    }
    ~StdoutSilencer()
    {
        fflush(stdout);
        std::cout.flush();
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};

RenderKernel make_kernel(RefScene* s, int w, int h, int spp, int bounces, Image& fb)
{
    return RenderKernel(w, h, spp, bounces, fb, s->triangles, s->materials, s->emissive,
                        s->material_indices, s->spheres, *s->bvh, s->env, s->env_cdf);
}

} // namespace

extern "C" {

int refo_max_threads() { return omp_get_max_threads(); }

void* refo_scene_from_obj(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) return nullptr;          // parse_obj would block on std::cin.get() then exit(1)
    fclose(f);

    ParsedOBJ parsed = Utils::parse_obj(path);
    RefScene* s = new RefScene;
    s->triangles = parsed.triangles;
    s->materials = parsed.materials;
    s->emissive = parsed.emissive_triangle_indices;
    s->material_indices = parsed.material_indices;
    s->build();
    return s;
}

// spheres4: center xyz + radius; sphere_prim: the primitive_index each sphere reports (sphere.h:11-53)
void* refo_scene_from_arrays(const float* tri9, int n_tri, const int* mat_idx,
                             const float* mats10, int n_mat,
                             const int* emissive, int n_emissive,
                             const float* spheres4, const int* sphere_prim, int n_sph,
                             const int* extra_mat_idx, int n_extra_mat_idx)
{
    RefScene* s = new RefScene;
    s->triangles.reserve(n_tri);
    for (int i = 0; i < n_tri; i++)
    {
        const float* p = tri9 + 9 * (size_t)i;
        s->triangles.push_back(Triangle(Point(p[0], p[1], p[2]), Point(p[3], p[4], p[5]), Point(p[6], p[7], p[8])));
    }
    s->material_indices.assign(mat_idx, mat_idx + n_tri);
    // material indices of analytic spheres live after the triangles' (main.cpp:19-31)
    for (int i = 0; i < n_extra_mat_idx; i++)
        s->material_indices.push_back(extra_mat_idx[i]);
    for (int i = 0; i < n_mat; i++)
    {
        const float* m = mats10 + 10 * (size_t)i;
        SimpleMaterial mat;
        mat.emission = Color(m[0], m[1], m[2], m[3]);
        mat.diffuse = Color(m[4], m[5], m[6], m[7]);
        mat.metalness = m[8];
        mat.roughness = m[9];
        s->materials.push_back(mat);
    }
    s->emissive.assign(emissive, emissive + n_emissive);
    for (int i = 0; i < n_sph; i++)
        s->spheres.push_back(Sphere(Point(spheres4[4 * i], spheres4[4 * i + 1], spheres4[4 * i + 2]), spheres4[4 * i + 3], sphere_prim[i]));
    s->build();
    return s;
}

void refo_scene_free(void* h) { delete (RefScene*)h; }

// out: n_tri, n_mat, n_emissive, n_material_indices, n_spheres, env_w, env_h
void refo_scene_counts(void* h, int* out7)
{
    RefScene* s = (RefScene*)h;
    out7[0] = (int)s->triangles.size();
    out7[1] = (int)s->materials.size();
    out7[2] = (int)s->emissive.size();
    out7[3] = (int)s->material_indices.size();
    out7[4] = (int)s->spheres.size();
    out7[5] = s->env.width();
    out7[6] = s->env.height();
}

double refo_scene_bvh_seconds(void* h) { return ((RefScene*)h)->bvh_build_seconds; }

void refo_scene_get(void* h, float* tri9, int* mat_idx, float* mats10, int* emissive)
{
    RefScene* s = (RefScene*)h;
    for (size_t i = 0; i < s->triangles.size(); i++)
        for (int v = 0; v < 3; v++)
        {
            tri9[9 * i + 3 * v + 0] = s->triangles[i][v].x;
            tri9[9 * i + 3 * v + 1] = s->triangles[i][v].y;
            tri9[9 * i + 3 * v + 2] = s->triangles[i][v].z;
        }
    for (size_t i = 0; i < s->material_indices.size(); i++)
        mat_idx[i] = s->material_indices[i];
    for (size_t i = 0; i < s->materials.size(); i++)
    {
        const SimpleMaterial& m = s->materials[i];
        float* o = mats10 + 10 * i;
        o[0] = m.emission.r; o[1] = m.emission.g; o[2] = m.emission.b; o[3] = m.emission.a;
        o[4] = m.diffuse.r; o[5] = m.diffuse.g; o[6] = m.diffuse.b; o[7] = m.diffuse.a;
        o[8] = m.metalness; o[9] = m.roughness;
    }
    for (size_t i = 0; i < s->emissive.size(); i++)
        emissive[i] = s->emissive[i];
}

// rgba: w*h*4 floats, row-major, row 0 first — stored verbatim into the reference Image
void refo_set_env(void* h, const float* rgba, int w, int height)
{
    RefScene* s = (RefScene*)h;
    s->env = Image(w, height);
    for (int i = 0; i < w * height; i++)
        s->env[i] = Color(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]);
    s->env_cdf = Utils::compute_env_map_cdf(s->env);
}

void refo_get_env_cdf(void* h, float* out)
{
    RefScene* s = (RefScene*)h;
    std::memcpy(out, s->env_cdf.data(), s->env_cdf.size() * sizeof(float));
}

// name ∈ {cornell, ganesha, ite_orb, dragon, mis}; out17 = 16 row-major view-matrix floats + fov_dist
int refo_camera_preset(const char* name, float* out17)
{
    std::string n(name);
    if (n == "cornell") camera_to17(Camera::CORNELL_BOX_CAMERA, out17);
    else if (n == "ganesha") camera_to17(Camera::GANESHA_CAMERA, out17);
    else if (n == "ite_orb") camera_to17(Camera::ITE_ORB_CAMERA, out17);
    else if (n == "dragon") camera_to17(Camera::PBRT_DRAGON_CAMERA, out17);
    else if (n == "mis") camera_to17(Camera::MIS_CAMERA, out17);
    else return -1;
    return 0;
}

// Camera(fov, RotationY(ry) * RotationX(rx) * Translation(tx,ty,tz)) — camera.h:27-32
void refo_camera_make(float fov_deg, float rot_x_deg, float rot_y_deg, float tx, float ty, float tz, float* out17)
{
    Camera cam(fov_deg, RotationY(rot_y_deg) * RotationX(rot_x_deg) * Translation(tx, ty, tz));
    camera_to17(cam, out17);
}

// rays6 out: origin xyz, direction xyz, via RenderKernel::get_camera_ray (render_kernel.cpp:56-73)
void refo_camera_rays(const float* cam17, int w, int h, const float* xy, int n, float* rays6)
{
    RefScene dummy;
    std::vector<Triangle> none;
    dummy.triangles = none;
    dummy.build();
    Image fb(1, 1);
    RenderKernel k = make_kernel(&dummy, w, h, 1, 1, fb);
    k.set_camera(camera_from17(cam17));
    for (int i = 0; i < n; i++)
    {
        Ray r = k.get_camera_ray(xy[2 * i], xy[2 * i + 1]);
        rays6[6 * i + 0] = r.origin.x; rays6[6 * i + 1] = r.origin.y; rays6[6 * i + 2] = r.origin.z;
        rays6[6 * i + 3] = r.direction.x; rays6[6 * i + 4] = r.direction.y; rays6[6 * i + 5] = r.direction.z;
    }
}

static void store_hit(bool found, const HitInfo& hi, int i, int* prim, float* t, float* extra8)
{
    prim[i] = found ? hi.primitive_index : -1;
    t[i] = found ? hi.t : -1.0f;
    if (extra8)
    {
        float* e = extra8 + 8 * (size_t)i;
        e[0] = hi.inter_point.x; e[1] = hi.inter_point.y; e[2] = hi.inter_point.z;
        e[3] = hi.normal_at_intersection.x; e[4] = hi.normal_at_intersection.y; e[5] = hi.normal_at_intersection.z;
        e[6] = hi.u; e[7] = hi.v;
    }
}

// mode 0: BVH::intersect (octree, triangles only)          bvh.cpp:62-65
// mode 1: RenderKernel::intersect_scene (brute force)       render_kernel.cpp:453-483
// mode 2: FlattenedBVH::intersect (prim index not reported) flattened_bvh.cpp:10-58
// mode 3: RenderKernel::intersect_scene_bvh (octree+spheres) render_kernel.cpp:485-502
static bool trace_one(RefScene* s, const RenderKernel& k, int mode, const Ray& ray, HitInfo& hi)
{
    switch (mode)
    {
    case 0: return s->bvh->intersect(ray, hi);
    case 1: return k.intersect_scene(ray, hi);
    case 2: return s->flat->intersect(ray, hi, s->triangles);
    default: return k.intersect_scene_bvh(ray, hi);
    }
}

int refo_trace(void* h, const float* rays6, int n, int mode, int* prim, float* t, float* extra8, int nthreads)
{
    RefScene* s = (RefScene*)h;
    if (mode == 2 && !s->flat)
        s->flat = new FlattenedBVH(s->bvh->flatten());
    Image fb(1, 1);
    RenderKernel k = make_kernel(s, 1, 1, 1, 1, fb);
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256) if (mode != 2)
    for (int i = 0; i < n; i++)
    {
        const float* r = rays6 + 6 * (size_t)i;
        Ray ray(Point(r[0], r[1], r[2]), Vector(r[3], r[4], r[5]));
        HitInfo hi;
        bool found = trace_one(s, k, mode, ray, hi);
        store_hit(found, hi, i, prim, t, extra8);
    }
    return 0;
}

// un-jittered primary rays: pixel (x,y) <-> get_camera_ray((float)x,(float)y); returns seconds of the ray loop
double refo_primary(void* h, const float* cam17, int w, int height, int mode, int* prim, float* t, int nthreads)
{
    RefScene* s = (RefScene*)h;
    Image fb(1, 1);
    RenderKernel k = make_kernel(s, w, height, 1, 1, fb);
    k.set_camera(camera_from17(cam17));
    if (nthreads > 0) omp_set_num_threads(nthreads);
    auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < height; y++)
        for (int x = 0; x < w; x++)
        {
            Ray ray = k.get_camera_ray((float)x, (float)y);
            HitInfo hi;
            bool found = trace_one(s, k, mode, ray, hi);
            store_hit(found, hi, y * w + x, prim, t, nullptr);
        }
    auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Full RenderKernel::render(); out_rgba = the tone-mapped framebuffer exactly as main.cpp sees it.
// Returns the wall seconds of render() alone (-1 on bad arguments).
double refo_render(void* h, const float* cam17, int w, int height, int spp, int bounces, float* out_rgba, int nthreads)
{
    RefScene* s = (RefScene*)h;
    if (s->env_cdf.empty() || height < 25) return -1.0;   // render_kernel.cpp:208 divides by height/25
    Image fb(w, height);
    RenderKernel k = make_kernel(s, w, height, spp, bounces, fb);
    k.set_camera(camera_from17(cam17));
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double seconds;
    {
        StdoutSilencer quiet;
        auto t0 = std::chrono::high_resolution_clock::now();
        k.render();
        auto t1 = std::chrono::high_resolution_clock::now();
        seconds = std::chrono::duration<double>(t1 - t0).count();
    }
    std::memcpy(out_rgba, fb.data(), sizeof(float) * 4 * (size_t)w * height);
    return seconds;
}

// Renders only the pixel rectangle [x0,x1)x[y0,y1) of a w x h frame with ray_trace_pixel (render_kernel.cpp:75-181):
// used to time / compare crops of frames too expensive to render in full on the CPU.
double refo_render_crop(void* h, const float* cam17, int w, int height, int spp, int bounces,
                        int x0, int y0, int x1, int y1, float* out_rgba_crop, int nthreads)
{
    RefScene* s = (RefScene*)h;
    if (s->env_cdf.empty()) return -1.0;
    Image fb(w, height);
    RenderKernel k = make_kernel(s, w, height, spp, bounces, fb);
    k.set_camera(camera_from17(cam17));
    if (nthreads > 0) omp_set_num_threads(nthreads);
    auto t0 = std::chrono::high_resolution_clock::now();
#pragma omp parallel for schedule(dynamic)
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++)
            k.ray_trace_pixel(x, y);
    auto t1 = std::chrono::high_resolution_clock::now();
    int cw = x1 - x0;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++)
        {
            Color c = fb[y * w + x];
            float* o = out_rgba_crop + 4 * ((size_t)(y - y0) * cw + (x - x0));
            o[0] = c.r; o[1] = c.g; o[2] = c.b; o[3] = c.a;
        }
    return std::chrono::duration<double>(t1 - t0).count();
}

// include/bvh_tests.h: 572 rays that must hit (+ expected points, tol 1e-5), 222 rays that must miss
void refo_golden_counts(int* n_hit, int* n_miss)
{
    *n_hit = (int)bvh_test_rays_inter.size();
    *n_miss = (int)bvh_test_rays_no_inter.size();
}

void refo_golden(float* hit_rays6, float* hit_points3, float* miss_rays6)
{
    for (size_t i = 0; i < bvh_test_rays_inter.size(); i++)
    {
        const Ray& r = bvh_test_rays_inter[i];
        const Point& p = bvh_test_rays_inter_result_points[i];
        float* o = hit_rays6 + 6 * i;
        o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z; o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z;
        hit_points3[3 * i] = p.x; hit_points3[3 * i + 1] = p.y; hit_points3[3 * i + 2] = p.z;
    }
    for (size_t i = 0; i < bvh_test_rays_no_inter.size(); i++)
    {
        const Ray& r = bvh_test_rays_no_inter[i];
        float* o = miss_rays6 + 6 * i;
        o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z; o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z;
    }
}

// xorshift32_generator known-answer hook (xorshift.h:10-31)
uint32_t refo_xorshift(uint32_t seed, int n_warmup, int n, float* out)
{
    xorshift32_generator g(seed);
    for (int i = 0; i < n_warmup; i++) g();
    uint32_t state_after_warmup = g.m_state.a;
    for (int i = 0; i < n; i++) out[i] = g();
    return state_after_warmup;
}

// the reference's own output stage: write_image_png (source/image_io.cpp:165-182) of a caller-supplied RGBA float image
int refo_write_png(const float* rgba, int w, int h, const char* path, int flip_y)
{
    Image img(w, h);
    for (int i = 0; i < w * h; i++) img[i] = Color(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]);
    return write_image_png(img, path, flip_y != 0) ? 0 : 1;
}

// the reference's own HDR round trip: write_image_hdr (source/image_io.cpp:208-215, stb's RGBE writer) and
// Utils::read_image_float (source/utils.cpp:100-124, stbi_loadf with the vertical flip): the ingest checker for b200rt_hdr_load
int refo_write_hdr(const float* rgba, int w, int h, const char* path, int flip_y)
{
    Image img(w, h);
    for (int i = 0; i < w * h; i++) img[i] = Color(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]);
    return write_image_hdr(img, path, flip_y != 0) ? 0 : 1;
}

// returns 0 and fills rgba_out (w*h*4, the Image read_image_float builds: alpha 0) when out_w/out_h match the file
int refo_read_image_float(const char* path, int flip_y, int expect_w, int expect_h, float* rgba_out)
{
    int w = 0, h = 0;
    Image img = Utils::read_image_float(path, w, h, flip_y != 0);
    if (w != expect_w || h != expect_h) return 1;
    for (int i = 0; i < w * h; i++) { Color c = img[i]; rgba_out[4 * i] = c.r; rgba_out[4 * i + 1] = c.g; rgba_out[4 * i + 2] = c.b; rgba_out[4 * i + 3] = c.a; }
    return 0;
}

// octree statistics (diagnostics only)
static void walk(const BVH::OctreeNode* n, int depth, long long* st)
{
    if (n->_is_leaf)
    {
        st[1]++;
        if (n->_triangles.empty()) st[2]++;
        st[3] = std::max<long long>(st[3], (long long)n->_triangles.size());
        st[4] = std::max<long long>(st[4], depth);
        return;
    }
    st[0]++;
    for (int i = 0; i < 8; i++) walk(n->_children[i], depth + 1, st);
}

// out5: inner nodes, leaves, empty leaves, max leaf size, max depth
void refo_octree_stats(void* h, long long* out5)
{
    for (int i = 0; i < 5; i++) out5[i] = 0;
    walk(((RefScene*)h)->bvh->_root, 0, out5);
}

} // extern "C"
