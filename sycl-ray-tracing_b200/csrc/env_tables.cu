// K5 env_tables (sm_100a): everything the integrator needs from the environment map is produced on the device.
//
//   k_env_expand_rgb      read_image_float's RGB -> RGBA expansion (source/utils.cpp:113-121: Color(r, g, b, 0) per texel)
//   k_env_luminance       Image::luminance_of_pixel (include/image.h:80-85; the weights are double constants)
//   k_env_cdf_serial      Utils::compute_env_map_cdf (source/utils.cpp:126-142): the running float sum IN THE REFERENCE'S ORDER.
//                         Float addition does not associate, and env_map_cdf_search (render_kernel.cpp:532-567) compares draws
//                         against these very values, so the table must be the serial sum bit for bit: a parallel scan would
//                         change which texel a draw selects. One CTA: the dependent chain of additions runs on one thread out
//                         of shared memory (4 cycles per texel) while the other threads stream the next chunk in and the last
//                         one out. 2 M texels: ~5 ms, one-shot.
//   k_env_row_cdf         the last column of the CDF, contiguous (the row search's probes)
//   alias table           Vose's table needs a sequential pairing of small and large cells; the same *kind* of table — every cell
//                         keeps probability q and defers 1 - q to exactly one other cell — is built here from prefix sums:
//                         lights (q < 1) in index order hand their deficit 1 - q to the heavy cell (q >= 1) whose cumulative
//                         excess interval contains the light's cumulative deficit; a heavy cell whose excess is used up keeps
//                         r = 1 + E_j - D(last light it served) and hands 1 - r to the next heavy cell. Three scans (cub, one-shot
//                         setup), two binary searches per texel, all in double. The implied distribution equals lum / total
//                         (tests compare it with the host's Vose table, b200rt_env_alias_table).
#include <cub/cub.cuh>

#include "kernels.h"

namespace b200rt {

__global__ void __launch_bounds__(256) k_env_expand_rgb(const float* __restrict__ rgb, size_t n, float4* __restrict__ rgba)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        rgba[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 0.0f);
}

__global__ void __launch_bounds__(256) k_env_luminance(const float4* __restrict__ env, size_t n, float* __restrict__ lum)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const float4 p = env[i];
        lum[i] = (float)(0.3086 * (double)p.x + 0.6094 * (double)p.y + 0.0820 * (double)p.z);
    }
}

constexpr int kCdfChunk = 4096;

__global__ void __launch_bounds__(256) k_env_cdf_serial(const float* __restrict__ lum, size_t n, float* __restrict__ cdf)
{
    __shared__ float buf[2][kCdfChunk];
    const size_t n_chunks = (n + kCdfChunk - 1) / kCdfChunk;
    float run = 0.0f;
    for (int i = threadIdx.x; i < kCdfChunk; i += blockDim.x) buf[0][i] = (size_t)i < n ? lum[i] : 0.0f;
    __syncthreads();
    for (size_t c = 0; c < n_chunks; c++)
    {
        float* cur = buf[c & 1];
        if (threadIdx.x == 0)
        {
            const int m = (int)min((size_t)kCdfChunk, n - c * kCdfChunk);
#pragma unroll 8
            for (int i = 0; i < m; i++) { run = run + cur[i]; cur[i] = run; }      // cdf[i] = cdf[i - 1] + lum[i], left to right
        }
        else if (c + 1 < n_chunks)
        {
            float* nxt = buf[(c + 1) & 1];
            const size_t base = (c + 1) * kCdfChunk;
            for (int i = threadIdx.x - 1; i < kCdfChunk; i += blockDim.x - 1) nxt[i] = base + i < n ? lum[base + i] : 0.0f;
        }
        __syncthreads();
        const size_t base = c * kCdfChunk;
        for (int i = threadIdx.x; i < kCdfChunk; i += blockDim.x) if (base + i < n) cdf[base + i] = cur[i];
        __syncthreads();
    }
}

__global__ void k_env_row_cdf(const float* __restrict__ cdf, int w, int h, float* __restrict__ row)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y < h) row[y] = cdf[(size_t)y * w + w - 1];
}

cudaError_t launch_env_expand_rgb(const float* rgb, size_t n, float4* rgba, cudaStream_t stream)
{
    if (!n) return cudaSuccess;
    k_env_expand_rgb<<<(unsigned int)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(rgb, n, rgba);
    return cudaGetLastError();
}

cudaError_t launch_env_luminance(const float4* env, size_t n, float* lum, cudaStream_t stream)
{
    if (!n) return cudaSuccess;
    k_env_luminance<<<(unsigned int)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(env, n, lum);
    return cudaGetLastError();
}

cudaError_t launch_env_cdf_serial(const float* lum, size_t n, float* cdf, cudaStream_t stream)
{
    if (!n) return cudaSuccess;
    k_env_cdf_serial<<<1, 256, 0, stream>>>(lum, n, cdf);
    return cudaGetLastError();
}

cudaError_t launch_env_row_cdf(const float* cdf, int w, int h, float* row, cudaStream_t stream)
{
    if (w <= 0 || h <= 0) return cudaSuccess;
    k_env_row_cdf<<<(h + 255) / 256, 256, 0, stream>>>(cdf, w, h, row);
    return cudaGetLastError();
}

// ---- guide table of the CDF search ---------------------------------------------------------------------------------------------------
// first index whose running sum exceeds v (n if none), v in double (the float table converts exactly)
__device__ __forceinline__ unsigned int cdf_first_above(const float* __restrict__ cdf, unsigned int n, double v)
{
    unsigned int lo = 0, hi = n;
    while (lo < hi) { const unsigned int m = (lo + hi) >> 1; if (v < (double)cdf[m]) hi = m; else lo = m + 1; }
    return lo;
}

// guide[b] = first_above((b / scale) * (1 - 1e-6)): a value v with (unsigned)(v * scale) == b (float multiply, |rounding| <= 2^-24)
// satisfies v >= (b / scale) * (1 - 2^-23) and v < ((b + 1) / scale) * (1 + 2^-23) < ((b + 2) / scale) * (1 - 1e-6), so its texel lies
// in [guide[b], guide[b + 2]] (first_above is non-decreasing in v)
__global__ void __launch_bounds__(256) k_env_cdf_guide(const float* __restrict__ cdf, unsigned int n, float scale, int n_buckets, unsigned int* __restrict__ guide)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_buckets + 1) return;
    guide[b] = b >= n_buckets ? n : cdf_first_above(cdf, n, ((double)b / (double)scale) * (1.0 - 1.0e-6));
}

__global__ void __launch_bounds__(256) k_env_cdf_monotone(const float* __restrict__ cdf, size_t n, int* __restrict__ violations)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (!(cdf[i] >= cdf[i - 1])) atomicAdd(violations, 1);
}

// builds the guide when the table allows it (finite positive total, non-decreasing sums); *built_out = 0 otherwise
cudaError_t build_env_cdf_guide(const float* cdf, size_t n, float total, int n_buckets, unsigned int* guide, float* scale_out, int* built_out, cudaStream_t stream)
{
    *built_out = 0;
    if (!(total > 0.0f) || !isfinite(total) || n < 2 || n >= 0x7fffffffu) return cudaSuccess;
    const float scale = (float)n_buckets / total;
    if (!(scale > 0.0f) || !isfinite(scale)) return cudaSuccess;
    int* d_viol = nullptr;
    cudaError_t e = cudaMalloc(&d_viol, sizeof(int));
    if (e != cudaSuccess) return e;
    int viol = 0;
    do
    {
        if ((e = cudaMemsetAsync(d_viol, 0, sizeof(int), stream)) != cudaSuccess) break;
        k_env_cdf_monotone<<<(unsigned int)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(cdf, n, d_viol);
        if ((e = cudaMemcpyAsync(&viol, d_viol, sizeof(int), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) break;
        if (viol) break;
        k_env_cdf_guide<<<(n_buckets + 2 + 255) / 256, 256, 0, stream>>>(cdf, (unsigned int)n, scale, n_buckets, guide);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        *scale_out = scale;
        *built_out = 1;
    } while (0);
    cudaFree(d_viol);
    return e;
}

// ---- alias table -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_alias_positive(const float* __restrict__ lum, size_t n, double* __restrict__ pos)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        pos[i] = lum[i] > 0.0f ? (double)lum[i] : 0.0;
}

// q = cell probability * n; lights carry their deficit, heavies their excess
__global__ void __launch_bounds__(256) k_alias_classify(const double* __restrict__ pos, size_t n, const double* __restrict__ total,
                                                        double* __restrict__ deficit, double* __restrict__ excess, int* __restrict__ light)
{
    const double scale = (double)n / *total;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const double q = pos[i] * scale;
        const bool l = q < 1.0;
        deficit[i] = l ? 1.0 - q : 0.0;
        excess[i] = l ? 0.0 : q - 1.0;
        light[i] = l ? 1 : 0;
    }
}

// D, E, R = inclusive scans of deficit, excess, light. Dl[k] = cumulative deficit through the first k lights (Dl[0] = 0),
// Eh[j] = cumulative excess through heavy j, Hidx[j] = texel of heavy j
__global__ void __launch_bounds__(256) k_alias_compact(const double* __restrict__ D, const double* __restrict__ E, const int* __restrict__ R,
                                                       const int* __restrict__ light, size_t n, double* __restrict__ Dl, double* __restrict__ Eh,
                                                       int* __restrict__ Hidx)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        if (i == 0) Dl[0] = 0.0;
        if (light[i]) Dl[R[i]] = D[i];
        else { const size_t j = i - (size_t)R[i]; Eh[j] = E[i]; Hidx[j] = (int)i; }
    }
}

__global__ void __launch_bounds__(256) k_alias_fill(const double* __restrict__ pos, const double* __restrict__ total, const int* __restrict__ R,
                                                    const int* __restrict__ light, size_t n, const double* __restrict__ Dl, const double* __restrict__ Eh,
                                                    const int* __restrict__ Hidx, float2* __restrict__ table)
{
    const int n_light = R[n - 1];
    const int n_heavy = (int)n - n_light;
    const double scale = (double)n / *total;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        float prob = 1.0f;
        int alias = (int)i;
        if (n_heavy > 0 && light[i])
        {
            // the heavy cell whose cumulative excess first reaches this light's cumulative deficit start: min { j : Eh[j] >= Dl[li] }
            const double d0 = Dl[R[i] - 1];
            int lo = 0, hi = n_heavy;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (Eh[m] >= d0) hi = m; else lo = m + 1; }
            prob = (float)(pos[i] * scale);
            alias = Hidx[min(lo, n_heavy - 1)];
        }
        else if (n_heavy > 0)
        {
            const int j = (int)(i - (size_t)R[i]);
            if (j + 1 < n_heavy)
            {
                // lights served by heavies 0..j: those whose cumulative deficit start is <= Eh[j]; k of them, through deficit Dl[k]
                const double ej = Eh[j];
                int lo = 0, hi = n_light;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (Dl[m] <= ej) lo = m + 1; else hi = m; }
                const double r = 1.0 + ej - Dl[lo];
                prob = (float)fmin(fmax(r, 0.0), 1.0);
                alias = Hidx[j + 1];
            }
        }
        table[i] = make_float2(prob, __int_as_float(alias));
    }
}

// builds the table of n cells from per-texel luminances; *total_out (device) receives the exact (double) luminance sum
cudaError_t build_env_alias_device(const float* lum, size_t n, float2* table, double* total_host, cudaStream_t stream)
{
    if (!n) return cudaErrorInvalidValue;
    cudaError_t e = cudaSuccess;
    double *pos = nullptr, *deficit = nullptr, *excess = nullptr, *Dl = nullptr, *Eh = nullptr, *total = nullptr;
    int *light = nullptr, *R = nullptr, *Hidx = nullptr;
    void* tmp = nullptr;
    const unsigned int grid = (unsigned int)std::min<size_t>((n + 255) / 256, 148 * 16);
    do
    {
        if ((e = cudaMalloc(&pos, n * sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMalloc(&deficit, n * sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMalloc(&excess, n * sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMalloc(&Dl, (n + 1) * sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMalloc(&Eh, n * sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMalloc(&total, sizeof(double))) != cudaSuccess) break;
        if ((e = cudaMalloc(&light, n * sizeof(int))) != cudaSuccess) break;
        if ((e = cudaMalloc(&R, n * sizeof(int))) != cudaSuccess) break;
        if ((e = cudaMalloc(&Hidx, n * sizeof(int))) != cudaSuccess) break;
        size_t need = 0, need2 = 0, need3 = 0;
        cub::DeviceReduce::Sum(nullptr, need, pos, total, (int)n, stream);
        cub::DeviceScan::InclusiveSum(nullptr, need2, deficit, deficit, (int)n, stream);
        cub::DeviceScan::InclusiveSum(nullptr, need3, light, R, (int)n, stream);
        need = std::max(need, std::max(need2, need3));
        if ((e = cudaMalloc(&tmp, std::max<size_t>(need, 16))) != cudaSuccess) break;
        k_alias_positive<<<grid, 256, 0, stream>>>(lum, n, pos);
        if ((e = cub::DeviceReduce::Sum(tmp, need, pos, total, (int)n, stream)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(total_host, total, sizeof(double), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) break;
        if (!(*total_host > 0.0)) { e = cudaErrorInvalidValue; break; }
        k_alias_classify<<<grid, 256, 0, stream>>>(pos, n, total, deficit, excess, light);
        if ((e = cub::DeviceScan::InclusiveSum(tmp, need, deficit, deficit, (int)n, stream)) != cudaSuccess) break;
        if ((e = cub::DeviceScan::InclusiveSum(tmp, need, excess, excess, (int)n, stream)) != cudaSuccess) break;
        if ((e = cub::DeviceScan::InclusiveSum(tmp, need, light, R, (int)n, stream)) != cudaSuccess) break;
        k_alias_compact<<<grid, 256, 0, stream>>>(deficit, excess, R, light, n, Dl, Eh, Hidx);
        k_alias_fill<<<grid, 256, 0, stream>>>(pos, total, R, light, n, Dl, Eh, Hidx, table);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        e = cudaStreamSynchronize(stream);
    } while (0);
    cudaFree(pos); cudaFree(deficit); cudaFree(excess); cudaFree(Dl); cudaFree(Eh); cudaFree(total); cudaFree(light); cudaFree(R); cudaFree(Hidx); cudaFree(tmp);
    return e;
}

} // namespace b200rt
