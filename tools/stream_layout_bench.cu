// Measurement tool (not part of the product): what HBM bandwidth does the ACCESS PATTERN of the step kernel
// reach on this GPU, independent of its arithmetic?  In-place update of NA fp64 arrays of N particles, one
// particle per thread, streaming loads/stores (ld.global.cs / st.global.cs) -- as
//   soa    : NA separate arrays (the product's layout: 10 read + 10 write streams per warp),
//   aosoa  : tiles of T particles, the NA fields of a tile contiguous (one 20 KB region per block),
//   copy   : one array copied to another (the 2-stream figure MEASURED_PEAKS.json quotes),
// plus `extra` write-only arrays (the deposit prepass: key + two colours).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_layout_bench stream_layout_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int NA = 10;
struct Ptrs { double *a[NA]; double *extra[3]; };

template <int EXTRA>
__global__ void __launch_bounds__(256, 4) soa_kernel(Ptrs p, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) v[k] = __ldcs(p.a[k] + i);
    double s = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k) { v[k] = v[k] * 1.0000001 + 1e-9; s += v[k]; }
#pragma unroll
    for (int k = 0; k < NA; ++k) __stcs(p.a[k] + i, v[k]);
#pragma unroll
    for (int k = 0; k < EXTRA; ++k) __stcs(p.extra[k] + i, s + k);
}

template <int EXTRA, int T>
__global__ void __launch_bounds__(256, 4) aosoa_kernel(double *base, Ptrs p, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *tile = base + (i / T) * (long)(T * NA) + (i % T);
    double v[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) v[k] = __ldcs(tile + k * T);
    double s = 0;
#pragma unroll
    for (int k = 0; k < NA; ++k) { v[k] = v[k] * 1.0000001 + 1e-9; s += v[k]; }
#pragma unroll
    for (int k = 0; k < NA; ++k) __stcs(tile + k * T, v[k]);
#pragma unroll
    for (int k = 0; k < EXTRA; ++k) __stcs(p.extra[k] + i, s + k);
}

__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ a, double2 *__restrict__ b, long n2)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) __stcs(b + i, __ldcs(a + i));
}

template <typename F>
static double time_ms(F f, int reps = 10)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int k = 0; k < 3; ++k) f();
    CK(cudaEventRecord(e0));
    for (int k = 0; k < reps; ++k) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / reps;
}

int main(int argc, char **argv)
{
    const long n = argc > 1 ? atol(argv[1]) : (1l << 26);
    Ptrs p;
    double *big;
    CK(cudaMalloc(&big, sizeof(double) * n * NA));
    CK(cudaMemset(big, 0, sizeof(double) * n * NA));
    for (int k = 0; k < NA; ++k) p.a[k] = big + (long)k * n;  // SoA view of the same memory
    for (int k = 0; k < 3; ++k) { CK(cudaMalloc(&p.extra[k], sizeof(double) * n)); }
    const int grid = (int)((n + 255) / 256);
    const double gb_state = 2.0 * NA * 8 * n / 1e9;
    double ms;
    ms = time_ms([&] { copy_kernel<<<(int)((n * NA / 2 / 2 + 255) / 256), 256>>>((const double2 *)big, (double2 *)(big + n * NA / 2), n * NA / 2 / 2); });
    printf("{\"pattern\": \"copy (2 streams, 128-bit)\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, 8.0 * n * NA / 1e9 / (ms * 1e-3));
    ms = time_ms([&] { soa_kernel<0><<<grid, 256>>>(p, n); });
    printf("{\"pattern\": \"soa 10r+10w\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, gb_state / (ms * 1e-3));
    ms = time_ms([&] { soa_kernel<3><<<grid, 256>>>(p, n); });
    printf("{\"pattern\": \"soa 10r+13w\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, (gb_state + 24.0 * n / 1e9) / (ms * 1e-3));
    ms = time_ms([&] { aosoa_kernel<0, 256><<<grid, 256>>>(big, p, n); });
    printf("{\"pattern\": \"aosoa256 10r+10w\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, gb_state / (ms * 1e-3));
    ms = time_ms([&] { aosoa_kernel<3, 256><<<grid, 256>>>(big, p, n); });
    printf("{\"pattern\": \"aosoa256 10r+13w\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, (gb_state + 24.0 * n / 1e9) / (ms * 1e-3));
    ms = time_ms([&] { aosoa_kernel<0, 32><<<grid, 256>>>(big, p, n); });
    printf("{\"pattern\": \"aosoa32 10r+10w\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, gb_state / (ms * 1e-3));
    return 0;
}
