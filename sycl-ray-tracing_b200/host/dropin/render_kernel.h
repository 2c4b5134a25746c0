// Drop-in replacement for the reference's include/render_kernel.h (TomClabault/SYCL-ray-tracing).
//
// Put this directory FIRST on the include path of the reference build and compile render_kernel_b200.cpp instead of
// source/render_kernel.cpp: source/main.cpp then drives the B200 path unchanged (main.cpp:94-113). Everything else
// (Triangle, Image, Camera, BVH, Utils, image_io ...) stays the reference's own code; only RenderKernel changes.
//
// Same constructor, set_camera() and render() as include/render_kernel.h:24-57. The object still owns none of the scene
// buffers (it stores the same references, :81-93); render() hands them to the C ABI of include/b200rt.h, which copies
// them to HBM, renders on the GPU(s) and writes the tone-mapped frame back into the caller's Image in place.
// ray_trace_pixel(x, y) (render_kernel.h:56, the DEBUG_PIXEL mode of render_kernel.cpp:186-197) renders one pixel on the GPU.
// The per-ray helper methods of the reference class (the BRDF and sampling functions) live on the device now
// (csrc/pt_device.cuh) and are not part of this host class.
// Environment: B200RT_GPUS=N renders every frame on N GPUs of this process (interleaved tiles, scene replicated, tiles gathered
// over NVLink on device 0); the image is bit-identical to the 1-GPU one.
#ifndef RENDER_KERNEL_H
#define RENDER_KERNEL_H

#include <vector>

#include "bvh.h"
#include "camera.h"
#include "color.h"
#include "image.h"
#include "simple_material.h"
#include "sphere.h"
#include "triangle.h"
#include "xorshift.h"

struct b200rt_scene;

class RenderKernel
{
public:
    RenderKernel(int width, int height, int render_samples, int max_bounces,
                 Image& image_buffer,
                 const std::vector<Triangle>& triangle_buffer_accessor,
                 const std::vector<SimpleMaterial>& materials_buffer_accessor,
                 const std::vector<int>& emissive_triangle_indices_buffer_accessor,
                 const std::vector<int>& materials_indices_buffer_accessor,
                 const std::vector<Sphere>& analytic_spheres_buffer,
                 BVH& bvh,
                 const Image& skysphere,
                 const std::vector<float>& env_map_cdf);
    ~RenderKernel();
    RenderKernel(RenderKernel&& other) noexcept;
    RenderKernel(const RenderKernel&) = delete;
    RenderKernel& operator=(const RenderKernel&) = delete;

    void set_camera(Camera camera) { m_camera = camera; }

    // render_kernel.cpp:56-73 (host restatement; handy for tests, not used by render())
    Ray get_camera_ray(float x, float y) const;

    // render_kernel.cpp:189-211: blocking; renders m_frame_buffer.width() x height() pixels in place
    void render();

    // render_kernel.cpp:75-181: one pixel, accumulated into and tone-mapped in the framebuffer in place like the reference's
    // (const there too: the framebuffer is a reference member)
    void ray_trace_pixel(int x, int y) const;

    // statistics of the last render() (rays = INTERSECT_SCENE-equivalent queries)
    unsigned long long last_rays() const { return m_last_rays; }
    double last_kernel_ms() const { return m_last_kernel_ms; }

private:
    int m_width, m_height;
    int m_render_samples;
    int m_max_bounces;

    Image& m_frame_buffer;

    const std::vector<Triangle>& m_triangle_buffer_access;
    const std::vector<SimpleMaterial>& m_materials_buffer_access;
    const std::vector<int>& m_emissive_triangle_indices_buffer;
    const std::vector<int>& m_materials_indices_buffer;

    const std::vector<Sphere>& m_sphere_buffer;

    const BVH& m_bvh;      // the reference octree is not traversed: the library builds its own flattened BVH from the triangles

    const Image& m_environment_map;
    const std::vector<float>& m_env_map_cdf;

    Camera m_camera;

    void ensure_scene() const;
    void fill_camera(float* camera17) const;

    mutable b200rt_scene* m_scene = nullptr;
    unsigned long long m_last_rays = 0;
    double m_last_kernel_ms = 0.0;
};

#endif
