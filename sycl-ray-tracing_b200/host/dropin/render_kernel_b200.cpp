// RenderKernel on the B200 path: replaces the reference's source/render_kernel.cpp in its build (see render_kernel.h here
// and INTEGRATION.md). Host code only marshals the reference's own containers into the C ABI of include/b200rt.h.
#include "render_kernel.h"

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "b200rt.h"

// the C ABI takes the reference's structs verbatim (include/b200rt.h "Layout contracts")
static_assert(std::is_trivially_copyable<Triangle>::value && sizeof(Triangle) == 36, "Triangle = 9 floats");
static_assert(std::is_trivially_copyable<SimpleMaterial>::value && sizeof(SimpleMaterial) == 40, "SimpleMaterial = 10 floats");
static_assert(std::is_trivially_copyable<Sphere>::value && sizeof(Sphere) == 20, "Sphere = 4 floats + int");
static_assert(sizeof(Color) == 16, "Color = 4 floats");

static void die(const char* what)
{
    // the reference's own error convention on this path is std::exit(1) (utils.cpp:26, :110)
    std::fprintf(stderr, "RenderKernel (b200rt): %s: %s\n", what, b200rt_last_error());
    std::exit(1);
}

RenderKernel::RenderKernel(int width, int height, int render_samples, int max_bounces, Image& image_buffer,
                           const std::vector<Triangle>& triangle_buffer_accessor,
                           const std::vector<SimpleMaterial>& materials_buffer_accessor,
                           const std::vector<int>& emissive_triangle_indices_buffer_accessor,
                           const std::vector<int>& materials_indices_buffer_accessor,
                           const std::vector<Sphere>& analytic_spheres_buffer, BVH& bvh, const Image& skysphere,
                           const std::vector<float>& env_map_cdf)
    : m_width(width), m_height(height), m_render_samples(render_samples), m_max_bounces(max_bounces),
      m_frame_buffer(image_buffer), m_triangle_buffer_access(triangle_buffer_accessor),
      m_materials_buffer_access(materials_buffer_accessor),
      m_emissive_triangle_indices_buffer(emissive_triangle_indices_buffer_accessor),
      m_materials_indices_buffer(materials_indices_buffer_accessor), m_sphere_buffer(analytic_spheres_buffer), m_bvh(bvh),
      m_environment_map(skysphere), m_env_map_cdf(env_map_cdf)
{
}

RenderKernel::RenderKernel(RenderKernel&& o) noexcept
    : m_width(o.m_width), m_height(o.m_height), m_render_samples(o.m_render_samples), m_max_bounces(o.m_max_bounces),
      m_frame_buffer(o.m_frame_buffer), m_triangle_buffer_access(o.m_triangle_buffer_access),
      m_materials_buffer_access(o.m_materials_buffer_access),
      m_emissive_triangle_indices_buffer(o.m_emissive_triangle_indices_buffer),
      m_materials_indices_buffer(o.m_materials_indices_buffer), m_sphere_buffer(o.m_sphere_buffer), m_bvh(o.m_bvh),
      m_environment_map(o.m_environment_map), m_env_map_cdf(o.m_env_map_cdf), m_camera(o.m_camera), m_scene(o.m_scene)
{
    o.m_scene = nullptr;
}

RenderKernel::~RenderKernel()
{
    if (m_scene) b200rt_scene_destroy(m_scene);
}

Ray RenderKernel::get_camera_ray(float x, float y) const
{
    float x_ndc_space = x / m_width * 2 - 1;
    x_ndc_space *= (float)m_width / m_height;
    float y_ndc_space = y / m_height * 2 - 1;
    Point origin = m_camera.view_matrix(Point(0.0f, 0.0f, 0.0f));
    Point target = m_camera.view_matrix(Point(x_ndc_space, y_ndc_space, m_camera.fov_dist));
    return Ray(origin, normalize(target - origin));
}

void RenderKernel::ensure_scene() const
{
    if (m_scene) return;
    // B200RT_GPUS=N: devices 0 .. N-1; B200RT_DEVICE_LIST=a,b,...: exactly these devices (a device may repeat)
    const char* gpus_env = std::getenv("B200RT_GPUS");
    int n_gpus = gpus_env ? std::atoi(gpus_env) : 1;
    std::vector<int> device_list;
    if (const char* list = std::getenv("B200RT_DEVICE_LIST"))
    {
        for (const char* p = list; *p;)
        {
            char* end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p) break;
            device_list.push_back((int)v);
            p = *end == ',' ? end + 1 : end;
        }
        if (!device_list.empty()) n_gpus = (int)device_list.size();
    }
    const float* env = m_environment_map.data();
    int rc = b200rt_scene_create_multi(
        reinterpret_cast<const float*>(m_triangle_buffer_access.data()), (int)m_triangle_buffer_access.size(),
        m_materials_indices_buffer.data(), (int)m_materials_indices_buffer.size(),
        reinterpret_cast<const float*>(m_materials_buffer_access.data()), (int)m_materials_buffer_access.size(),
        m_emissive_triangle_indices_buffer.data(), (int)m_emissive_triangle_indices_buffer.size(),
        m_sphere_buffer.empty() ? nullptr : m_sphere_buffer.data(), (int)m_sphere_buffer.size(),
        env, 4, m_environment_map.width(), m_environment_map.height(), m_env_map_cdf.data(),
        nullptr /* build the flattened BVH from the triangles */,
        device_list.empty() ? nullptr /* devices 0 .. n-1 (one device: the current one) */ : device_list.data(),
        n_gpus > 1 ? n_gpus : 1, &m_scene);
    if (rc) die("scene upload failed");
}

void RenderKernel::fill_camera(float* camera17) const
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            camera17[4 * i + j] = m_camera.view_matrix.m[i][j];
    camera17[16] = m_camera.fov_dist;
}

void RenderKernel::render()
{
    ensure_scene();
    float camera17[17];
    fill_camera(camera17);

    b200rt_render_options opts;
    b200rt_default_render_options(&opts);
    b200rt_stats stats;
    // like the reference, iterate the framebuffer's own size (render_kernel.cpp:199-201); the framebuffer is read,
    // accumulated into and tone-mapped in place (:167-180)
    int rc = b200rt_render(m_scene, camera17, m_frame_buffer.width(), m_frame_buffer.height(), m_render_samples, m_max_bounces,
                           m_frame_buffer.data(), &opts, &stats);
    if (rc) die("render failed");
    m_last_rays = stats.rays;
    m_last_kernel_ms = stats.kernel_ms;
}

void RenderKernel::ray_trace_pixel(int x, int y) const
{
    ensure_scene();
    float camera17[17];
    fill_camera(camera17);
    // the pixel's mean radiance tone-mapped on a black framebuffer comes back from the device; the reference adds the
    // mean to the pixel's current value first (:169). A pixel is rendered once, from Color::Black(), in every use the reference makes
    // of this method (render(), DEBUG_PIXEL), which is the case reproduced bit for bit here.
    float rgba[4];
    int rc = b200rt_render_region(m_scene, camera17, m_frame_buffer.width(), m_frame_buffer.height(), m_render_samples, m_max_bounces,
                                  x, y, x + 1, y + 1, rgba, nullptr, nullptr);
    if (rc) die("ray_trace_pixel failed");
    Color& px = m_frame_buffer[y * m_frame_buffer.width() + x];
    px = Color(rgba[0], rgba[1], rgba[2], rgba[3]);
}
