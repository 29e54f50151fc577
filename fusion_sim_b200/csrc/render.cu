// render.cu -- out.canvas: what the page draws each frame (sm_100a).
//
// programBMag (empic.js:479-482) drawn to the RGBA8 canvas without blending, then
// programDensity (empic.js:1101-1105) blended SRC_ALPHA,ONE (out.density :1497-1504).  The canvas
// is a fixed-point target: each draw's colour is clamped to [0,1] and stored as round(255 c);
// NaN converts to 0.  Rows are written in canvas order (top row = GL row nz-1).
#include <algorithm>

#include "common.cuh"

namespace fsim {

template <typename Real>
__device__ __forceinline__ Real clamp01(Real v)
{
    if (!(v > (Real)0)) return (Real)0;
    if (v > (Real)1) return (Real)1;
    return v;
}
template <typename Real>
__device__ __forceinline__ Real quant8(Real v)
{
    return ffloor(v * (Real)255.0 + (Real)0.5);
}

// first draw, programBMag (empic.js:479-482): the colour of |B| and its direction, as the 8-bit canvas stores it.
// B is static between field changes, so this image is drawn once and kept (fsim_sim::bmag) until B changes.
template <typename Real>
__global__ void __launch_bounds__(256)
bmag_kernel(const Real *__restrict__ B, uchar4 *__restrict__ base, int nr, int row0, int own0, int own_rows)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nr * own_rows) return;
    const int i = (int)(t % nr), j = own0 + (int)(t / nr);
    const size_t c = (size_t)i + (size_t)(j - row0) * nr;
    const Real Bx = B[3 * c], By = B[3 * c + 1], Bz = B[3 * c + 2];
    const Real mag = fsqrt(Bx * Bx + By * By + Bz * Bz);
    const Real dx = Bx / mag, dz = Bz / mag;
    const Real mn = (dz < (Real)0) ? dz : (Real)0;
    const Real mx = (dz > (Real)0) ? dz : (Real)0;
    Real c1[4];
    c1[0] = mag * ((mn < (Real)0) ? -mn : mn);
    c1[1] = mag * dx;
    c1[2] = mag * ((mx < (Real)0) ? -mx : mx);
    c1[3] = (Real)1.0;
    uchar4 o;
    uint8_t *po = reinterpret_cast<uint8_t *>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) po[q] = (uint8_t)quant8(clamp01(c1[q]));
    base[t] = o;
}

// second draw, programDensity (empic.js:1101-1105) blended SRC_ALPHA, ONE over the first (out.density :1497-1504)
template <typename Real>
__global__ void __launch_bounds__(256)
render_kernel(const uchar4 *__restrict__ base, const Real *__restrict__ avg_alpha, int pitch,
              uint8_t *__restrict__ rgba, int nr, int nz, int row0, int own0, int own_rows)
{
    // v / 255 for the 256 values a stored canvas byte can take: one IEEE division per 256-cell group of a block's
    // grid-stride walk instead of four per cell (the table is built once per block)
    __shared__ Real inv255[256];
    inv255[threadIdx.x] = (Real)threadIdx.x / (Real)255.0;  // blockDim.x == 256
    __syncthreads();
    // the alpha channel of both draws is constant: src = (d, d, d, 1) * RENDER_DENSITY
    const Real sa = clamp01((Real)FSIM_RENDER_DENSITY * (Real)1.0);
    const uint32_t total = (uint32_t)nr * (uint32_t)own_rows;  // cells fit 31 bits (the sort key does)
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int i = (int)(t % (uint32_t)nr), j = own0 + (int)(t / (uint32_t)nr);
        const uchar4 b8 = base[t];
        const uint8_t *pb = reinterpret_cast<const uint8_t *>(&b8);
        const Real a = avg_alpha[(size_t)(j - row0) * pitch + i];  // alpha plane of the running average
        const Real sc = (Real)FSIM_RENDER_DENSITY * a;
        const Real src[4] = {sc, sc, sc, (Real)FSIM_RENDER_DENSITY * (Real)1.0};
        uchar4 o;
        uint8_t *po = reinterpret_cast<uint8_t *>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const Real dst = inv255[pb[q]];
            const Real out = clamp01(src[q]) * sa + dst;
            po[q] = (uint8_t)quant8(clamp01(out));
        }
        reinterpret_cast<uchar4 *>(rgba)[(size_t)i + (size_t)(nz - 1 - j) * nr] = o;
    }
}

int launch_render(fsim_sim *s, uint8_t *dev_rgba, cudaStream_t st)
{
    const int64_t nown = (int64_t)s->nr * s->own_rows;
    if (!s->bmag) FSIM_CUDA(cudaMalloc((void **)&s->bmag, sizeof(uchar4) * (size_t)std::max<int64_t>(nown, 1)));
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        if (!s->bmag_valid) {  // B changed (or first image): redraw the |B| layer -- on the stream the field builders use
            bmag_kernel<Real><<<grid_for(nown, 256), 256, 0, st>>>((const Real *)s->B, (uchar4 *)s->bmag, s->nr, s->row0, s->own0,
                                                                   s->own_rows);
            FSIM_CUDA(cudaGetLastError());
            s->launches++;
            s->bmag_valid = true;
        }
        Bracket b(s, "render", st);
        const int64_t blocks = std::min<int64_t>((nown + 255) / 256, (int64_t)s->nsm * 8);
        render_kernel<Real><<<(unsigned)std::max<int64_t>(blocks, 1), 256, 0, st>>>(
            (const uchar4 *)s->bmag, (const Real *)s->avg + 3 * s->plane, s->pitch, dev_rgba, s->nr, s->nz, s->row0,
            s->own0, s->own_rows);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

}  // namespace fsim
