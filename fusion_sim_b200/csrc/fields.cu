// fields.cu -- per-cell maps: precalc() and the static-field builders (sm_100a).
//
// precalc  : programPre1/2/3 + programPreA (empic.js:506-659, out.precalc :1413-1434) fused into
//            one pass: (E,B) -> the 8-real cell record (B, f, 1 - h^2|B|^2 f, A), see common.cuh.
// add_loop : out.addCurrentLoop (empic.js:1352-1363).  The reference renders two Biot-Savart
//            tables (programCurrentLoopShape :308-326, u_R = 0.5 and 0.1) and samples them NEAREST
//            at a scaled coordinate (programCurrentLoop :367-377).  The sampled texel value is a
//            pure function of the texel index, so it is evaluated on the fly for exactly the
//            texel the reference would fetch: same arithmetic, no 2 x nr x nz tables held in HBM
//            and no dependence on which slab of the grid a GPU owns.
// add_uniform : addCurrentZ :404, addBZ :429, addBTheta :454.
#include "common.cuh"

namespace fsim {

__constant__ double c_cos_f64[FSIM_NQUAD];
__constant__ float c_cos_f32[FSIM_NQUAD];
template <typename Real> __device__ __forceinline__ Real cos_tab(int k);
template <> __device__ __forceinline__ double cos_tab<double>(int k) { return c_cos_f64[k]; }
template <> __device__ __forceinline__ float cos_tab<float>(int k) { return c_cos_f32[k]; }

int upload_costab(const double *c64, const float *c32)
{
    FSIM_CUDA(cudaMemcpyToSymbol(c_cos_f64, c64, sizeof(double) * FSIM_NQUAD));
    FSIM_CUDA(cudaMemcpyToSymbol(c_cos_f32, c32, sizeof(float) * FSIM_NQUAD));
    return FSIM_OK;
}

// the 8-real record leaves as whole 32-byte sectors: two 256-bit stores in fp64 (st.global.v4.f64, sm_100), two 128-bit
// stores in fp32 -- eight scalar stores at a 64-byte stride wrote every sector of the warp's 2 KB four (fp32: eight)
// times over and kept the kernel at 0.5 of the copy bandwidth on L1TEX wavefronts
__device__ __forceinline__ void st_rec(double *p, const double (&o)[RECSTRIDE])
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(o[0]), "d"(o[1]), "d"(o[2]), "d"(o[3]) : "memory");
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "d"(o[4]), "d"(o[5]), "d"(o[6]), "d"(o[7]) : "memory");
}
__device__ __forceinline__ void st_rec(float *p, const float (&o)[RECSTRIDE])
{
    reinterpret_cast<float4 *>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4 *>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
}

template <typename Real>
__global__ void __launch_bounds__(256)
precalc_kernel(const Real *__restrict__ E, const Real *__restrict__ B, Real *__restrict__ rec,
               int64_t ncell, Real h, Real kr, Real kz, int corrected)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const Real Bx = B[3 * c], By = B[3 * c + 1], Bz = B[3 * c + 2];
    const Real Ex = E[3 * c], Ey = E[3 * c + 1], Ez = E[3 * c + 2];
    const Real Bmag = fsqrt(Bx * Bx + By * By + Bz * Bz);  // length(B)
    const Real hB2 = h * h * Bmag * Bmag;
    const Real f = (Real)2.0 / ((Real)1.0 + hB2);
    const Real one_m = (Real)1.0 - hB2 * f;
    Real o[RECSTRIDE];
    // what programPre1/2/3 need of B (empic.js:520-522); the nine entries are rebuilt by boris_rows()
    o[REC_BX] = Bx; o[REC_BY] = By; o[REC_BZ] = Bz;
    o[REC_F] = f;
    o[REC_ONEM] = one_m;
    // programPreA, empic.js:645-647
    const Real cx = Ey * Bz - Ez * By;
    const Real cy = Ez * Bx - Ex * Bz;
    const Real cz = Ex * By - Ey * Bx;
    const Real d = Ex * Bx + Ey * By + Ez * Bz;
    const Real t1 = h * ((Real)2.0 - hB2 * f);
    const Real t2 = h * h * f;
    Real ax, ay, az;
    if (corrected) {  // FSIM_FLAG_CORRECTED_PREA: textbook h (E.B) B
        ax = (t1 * Ex + t2 * (cx + h * d * Bx)) / (Real)FSIM_C_LIGHT;
        ay = (t1 * Ey + t2 * (cy + h * d * By)) / (Real)FSIM_C_LIGHT;
        az = (t1 * Ez + t2 * (cz + h * d * Bz)) / (Real)FSIM_C_LIGHT;
    } else {  // as written in the reference: scalar u_h*dot(E,B) added to each component
        const Real hd = h * d;
        ax = (t1 * Ex + t2 * (cx + hd)) / (Real)FSIM_C_LIGHT;
        ay = (t1 * Ey + t2 * (cy + hd)) / (Real)FSIM_C_LIGHT;
        az = (t1 * Ez + t2 * (cz + hd)) / (Real)FSIM_C_LIGHT;
    }
    o[REC_AX] = ax * kr;
    o[REC_AY] = ay * kr;
    o[REC_AZ] = az * kz;
    st_rec(rec + RECSTRIDE * c, o);
}

// accessor support: expand the records to the reference's four textures R1 R2 R3 A (12 reals)
template <typename Real>
__global__ void __launch_bounds__(256)
expand_records_kernel(const Real *__restrict__ rec, double *__restrict__ out, int64_t ncell, Real h, Real k13,
                      Real k31)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    Real R[9];
    boris_rows<Real>(rec + RECSTRIDE * c, h, k13, k31, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) out[FSIM_CELLREC * c + k] = (double)R[k];
    out[FSIM_CELLREC * c + 9] = (double)rec[RECSTRIDE * c + REC_AX];
    out[FSIM_CELLREC * c + 10] = (double)rec[RECSTRIDE * c + REC_AY];
    out[FSIM_CELLREC * c + 11] = (double)rec[RECSTRIDE * c + REC_AZ];
}

int launch_expand_records(fsim_sim *s, double *dev_out)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        expand_records_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>(
            (const Real *)s->cellrec, dev_out, s->ncell_local, (Real)s->h, (Real)s->k13, (Real)s->k31);
        FSIM_CUDA(cudaGetLastError());
        s->launches++;
        return (int)FSIM_OK;
    });
}

int launch_precalc(fsim_sim *s)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        Bracket b(s, "precalc");
        precalc_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>(
            (const Real *)s->E, (const Real *)s->B, (Real *)s->cellrec, s->ncell_local, (Real)s->h,
            (Real)s->kr, (Real)s->kz,
            (s->spec.flags & FSIM_FLAG_CORRECTED_PREA) ? 1 : 0);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

template <typename Real>
__global__ void __launch_bounds__(128)
add_loop_kernel(Real *__restrict__ B, int nr, int nz, int row0, int rows, Real R, Real Z, Real I)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    const int i = (int)(c % nr);
    int j = (int)(c / nr) + row0;
    if (j < 0) j += nz;  // periodic z: a ghost row holds the field of the row it mirrors
    if (j >= nz) j -= nz;
    const Real u = ((Real)i + (Real)0.5) / (Real)nr;
    const Real v = ((Real)j + (Real)0.5) / (Real)nz;
    // programCurrentLoop, empic.js:367-377
    const Real a = u / R;
    const Real b = (v - Z) / R;
    const Real sgn = (b > (Real)0) ? (Real)1 : ((b < (Real)0) ? (Real)-1 : (Real)0);
    const Real ab = (b < (Real)0) ? -b : b;
    Real Rt;
    int ti, tj;
    if (a > (Real)FSIM_LOOP_FAR || b > (Real)FSIM_LOOP_FAR) {
        Rt = (Real)0.1;  // u_shape_tenth, empic.js:341
        ti = tex_idx(a / (Real)10.0, nr);
        tj = tex_idx(ab / (Real)10.0, nz);
    } else {
        Rt = (Real)0.5;  // u_shape_half, empic.js:336
        ti = tex_idx(a / (Real)2.0, nr);
        tj = tex_idx(ab / (Real)2.0, nz);
    }
    // programCurrentLoopShape evaluated at texel (ti,tj), empic.js:308-326
    const Real tu = ((Real)ti + (Real)0.5) / (Real)nr;
    const Real tv = ((Real)tj + (Real)0.5) / (Real)nz;
    const Real constant = Rt * (Real)FSIM_QUAD_SCALE * (Real)FSIM_MU0 / ((Real)4.0 * (Real)FSIM_PI_GLSL);
    Real Bx = (Real)0, Bz = (Real)0;
    for (int k = 0; k < FSIM_NQUAD; ++k) {
        const Real cosine = cos_tab<Real>(k);
        const Real r = fsqrt(Rt * Rt + tu * tu + tv * tv - (Real)2.0 * tu * Rt * cosine);
        const Real factor = (r > (Real)0) ? constant * (Real)1.0 / (r * r * r) : (Real)0;
        Bx += tv * factor * cosine;
        Bz += factor * (Rt - tu * cosine);
    }
    // field = u_I * vec4(sign(b),1,1,1) * texel, blended ONE,ONE (table .y is 0)
    Real *o = B + 3 * c;
    o[0] = o[0] + (I * sgn) * Bx;
    o[1] = o[1] + (I * (Real)1.0) * (Real)0.0;
    o[2] = o[2] + (I * (Real)1.0) * Bz;
}

int launch_add_loop(fsim_sim *s, double R, double Z, double I)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        Bracket b(s, "add_loop");
        add_loop_kernel<Real><<<grid_for(s->ncell_local, 128), 128, 0, s->stream>>>(
            (Real *)s->B, s->nr, s->nz, s->row0, s->rows, (Real)R, (Real)Z, (Real)I);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

template <typename Real>
__global__ void __launch_bounds__(256)
add_uniform_kernel(Real *__restrict__ B, int nr, int64_t ncell, int kind, Real val)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int i = (int)(c % nr);
    const Real u = ((Real)i + (Real)0.5) / (Real)nr;
    Real *o = B + 3 * c;
    if (kind == 0)  // addCurrentZ, empic.js:404
        o[1] = o[1] + val * (Real)FSIM_MU0 / ((Real)2.0 * (Real)FSIM_PI_GLSL * u);
    else if (kind == 1)  // addBZ, empic.js:429
        o[2] = o[2] + val;
    else  // addBTheta, empic.js:454
        o[1] = o[1] + val;
}

int launch_add_uniform(fsim_sim *s, int kind, double val)
{
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        Bracket b(s, "add_uniform");
        add_uniform_kernel<Real><<<grid_for(s->ncell_local, 256), 256, 0, s->stream>>>(
            (Real *)s->B, s->nr, s->ncell_local, kind, (Real)val);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
}

}  // namespace fsim
