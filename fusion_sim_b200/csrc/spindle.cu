// spindle.cu -- out.addSpindleCuspPlasmaField (empic.js:1369): a spindle cusp with a field-excluding plasma.
//
// "Next" row N3 of SURVEY.md section 8f, second half.  spindle.makeSpindleCuspPlasmaField does not run in the
// reference (spindle.js:57,328,333,624,643,651: undefined names, a shader without its semicolon, unit slips), so
// this is built from its INTENT -- spindle.js:27-30 "solves the boundary conditions for a perfect conductor in
// center of a spindle cusp magnetic field", :138-176 the surface arc and its loop pairs, :418-630 the normal field
// of fixed and variable loops at the surface points, :632-654 solve A x = b with makeSORIterative and superpose
// the loops -- on the reference's own solver, fsim_jacobi_* (jacobi.cu).  The specification is written out in
// include/fusionsim.h; the test suite holds a second, CPU statement of it and the two must agree bit for
// bit.  PARITY UNPINNED (there is nothing in the reference to be equal to).
//
//   1. surface geometry on the host (libm, fp64);
//   2. node_field_kernel: the field of every surface loop (and its mirror image) at every collocation point --
//      (L+1) x L x 2 quadratures of 1000 terms, one thread each;  matrix_kernel: A, rhs;
//   3. gauge + weighted-Jacobi solve (fsim_jacobi_*, the device routine of matrix_webgl.makeSORIterative);
//   4. add_loops_kernel: coils + surface loops superposed on B, one thread per cell.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace fsim {

// cosines of the quadrature angles (empic.js:317), host libm, uploaded at the first solve (constant memory is
// per translation unit: fields.cu holds its own copy of the same table)
__constant__ double c_cos_f64[FSIM_NQUAD];
__constant__ float c_cos_f32[FSIM_NQUAD];

// field (tesla per ampere) of a loop of radius Rl at height Zl at the point (x, z): the quadrature of
// programCurrentLoopShape (empic.js:308-326) at the exact relative position, midpoint-rule weight 2 pi / 1000
template <typename Real>
__device__ __forceinline__ void loop_field(Real Rl, Real Zl, Real x, Real z, Real &br, Real &bz)
{
    const Real dz = z - Zl;
    const Real K = Rl * (Real)FSIM_SPINDLE_QW * (Real)FSIM_MU0 / ((Real)4.0 * (Real)FSIM_PI_GLSL);
    Real Br = (Real)0, Bz = (Real)0;
    for (int k = 0; k < FSIM_NQUAD; ++k) {
        Real c;
        if constexpr (sizeof(Real) == 8) c = c_cos_f64[k];
        else c = c_cos_f32[k];
        const Real rho = fsqrt(Rl * Rl + x * x + dz * dz - (Real)2.0 * x * Rl * c);
        const Real f = (rho > (Real)0) ? K / (rho * rho * rho) : (Real)0;
        Br += dz * f * c;
        Bz += f * (Rl - x * c);
    }
    br = Br;
    bz = Bz;
}

// nf[node][point] = field of the loop through `node` minus that of its mirror image about z = height/2
__global__ void __launch_bounds__(128)
node_field_kernel(const double *__restrict__ nodes, const double *__restrict__ points, double height, double *__restrict__ nf)
{
    constexpr int L = FSIM_SPINDLE_L;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (L + 1) * L) return;
    const int l = t / L, p = t % L;
    double r0, z0, r1, z1;
    loop_field<double>(nodes[2 * l], nodes[2 * l + 1], points[2 * p], points[2 * p + 1], r0, z0);
    loop_field<double>(nodes[2 * l], height - nodes[2 * l + 1], points[2 * p], points[2 * p + 1], r1, z1);
    nf[2 * t] = r0 - r1;
    nf[2 * t + 1] = z0 - z1;
}

// A[p][e] = n_p . (field of element e = loop +1 through node e, loop -1 through node e+1), rhs[p] = -n_p . (coil field);
// gauge x_{L-1} = 0: row and column L-1 are those of the identity
__global__ void __launch_bounds__(256)
spindle_matrix_kernel(const double *__restrict__ nf, const double *__restrict__ points, const double *__restrict__ normals,
                      double height, double coil_r, double coil_I, double *__restrict__ A, double *__restrict__ rhs)
{
    constexpr int L = FSIM_SPINDLE_L;
    const int p = blockIdx.x, e = threadIdx.x;  // <<<L, L>>>
    const double nx = normals[2 * p], nz = normals[2 * p + 1];
    const double er = nf[2 * (e * L + p)] - nf[2 * ((e + 1) * L + p)];
    const double ez = nf[2 * (e * L + p) + 1] - nf[2 * ((e + 1) * L + p) + 1];
    double a = nx * er + nz * ez;
    if (p == L - 1 || e == L - 1) a = (p == e) ? 1.0 : 0.0;
    A[p * L + e] = a;
    if (e == 0) {
        double r0, z0, r1, z1;
        loop_field<double>(coil_r, 0.0, points[2 * p], points[2 * p + 1], r0, z0);
        loop_field<double>(coil_r, height, points[2 * p], points[2 * p + 1], r1, z1);
        rhs[p] = (p == L - 1) ? 0.0 : -(coil_I * (nx * (r0 - r1) + nz * (z0 - z1)));
    }
}

// B += sum of I_l x (loop field at the cell centre), loops in the order given, in the working precision
template <typename Real>
__global__ void __launch_bounds__(128)
add_loops_kernel(Real *__restrict__ B, int nr, int nz, int row0, int rows, Real dr, Real dzc, int nloops,
                 const double *__restrict__ loops)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)nr * rows) return;
    const int i = (int)(c % nr);
    int j = (int)(c / nr) + row0;
    if (j < 0) j += nz;  // periodic z: ghost rows mirror the wrapped rows
    if (j >= nz) j -= nz;
    const Real x = ((Real)i + (Real)0.5) * dr;
    const Real z = ((Real)j + (Real)0.5) * dzc;
    Real sr = (Real)0, sz = (Real)0;
    for (int l = 0; l < nloops; ++l) {
        Real br, bz;
        loop_field<Real>((Real)loops[3 * l], (Real)loops[3 * l + 1], x, z, br, bz);
        const Real I = (Real)loops[3 * l + 2];
        sr += I * br;
        sz += I * bz;
    }
    Real *o = B + 3 * c;
    o[0] = o[0] + sr;
    o[2] = o[2] + sz;
}

static void spindle_geometry(double radius, double height, std::vector<double> &nodes, std::vector<double> &points,
                             std::vector<double> &normals)
{
    constexpr int L = FSIM_SPINDLE_L;
    const double a = FSIM_SPINDLE_A;
    const double s = height / (2.0 * radius);
    const double R = radius * sqrt(1.0 + a * a);          // spindle.js:139
    const double alpha = atan(a);                         // :141
    const double theta = alpha + FSIM_PI;                 // :144
    const double arc = 0.5 * FSIM_PI - 2.0 * alpha;       // :146
    nodes.resize(2 * (L + 1)); points.resize(2 * L); normals.resize(2 * L);
    for (int l = 0; l <= L; ++l) {                        // :161-168 (arc / L where the reference has arc / 1000)
        const double phi = (double)l * arc / (double)L + theta;
        nodes[2 * l] = (l == 0) ? 0.0 : R * cos(-phi) + radius;  // node 0 lies on the axis
        nodes[2 * l + 1] = s * (R * sin(-phi));
    }
    for (int p = 0; p < L; ++p) {                         // :148-160
        const double phi = ((double)p + 0.5) * arc / (double)L + theta;
        points[2 * p] = R * cos(-phi) + radius;
        points[2 * p + 1] = s * (R * sin(-phi));
        const double nx = -(s * cos(-phi)), nz = -sin(-phi);
        const double len = sqrt(nx * nx + nz * nz);
        normals[2 * p] = nx / len;
        normals[2 * p + 1] = nz / len;
    }
}

int spindle_solve(fsim_sim *s, double coil_r, double B_c, double beta_c)
{
    constexpr int L = FSIM_SPINDLE_L;
    const double radius = s->spec.radius, height = s->spec.height;
    std::vector<double> nodes, points, normals;
    spindle_geometry(radius, height, nodes, points, normals);
    const double coil_I = 2.0 * coil_r * B_c / FSIM_MU0;
    {
        double c64[FSIM_NQUAD];
        float c32[FSIM_NQUAD];
        host_cos_tables(c64, c32);
        FSIM_CUDA(cudaMemcpyToSymbolAsync(c_cos_f64, c64, sizeof c64, 0, cudaMemcpyHostToDevice, s->stream));
        FSIM_CUDA(cudaMemcpyToSymbolAsync(c_cos_f32, c32, sizeof c32, 0, cudaMemcpyHostToDevice, s->stream));
        FSIM_CUDA(cudaStreamSynchronize(s->stream));  // the tables live on this stack frame
    }

    // device scratch: nodes | points | normals | nf | A | rhs
    const size_t n_nodes = 2 * (L + 1), n_pts = 2 * L, n_nf = 2 * (size_t)(L + 1) * L, n_A = (size_t)L * L;
    const size_t total = n_nodes + 2 * n_pts + n_nf + n_A + L;
    FSIM_TRY(ensure_stage(s, sizeof(double) * total));
    double *d_nodes = (double *)s->stage, *d_pts = d_nodes + n_nodes, *d_nrm = d_pts + n_pts, *d_nf = d_nrm + n_pts,
           *d_A = d_nf + n_nf, *d_rhs = d_A + n_A;
    FSIM_CUDA(cudaMemcpyAsync(d_nodes, nodes.data(), sizeof(double) * n_nodes, cudaMemcpyHostToDevice, s->stream));
    FSIM_CUDA(cudaMemcpyAsync(d_pts, points.data(), sizeof(double) * n_pts, cudaMemcpyHostToDevice, s->stream));
    FSIM_CUDA(cudaMemcpyAsync(d_nrm, normals.data(), sizeof(double) * n_pts, cudaMemcpyHostToDevice, s->stream));
    node_field_kernel<<<grid_for((L + 1) * L, 128), 128, 0, s->stream>>>(d_nodes, d_pts, height, d_nf);
    spindle_matrix_kernel<<<L, L, 0, s->stream>>>(d_nf, d_pts, d_nrm, height, coil_r, coil_I, d_A, d_rhs);
    FSIM_CUDA(cudaGetLastError());
    s->launches += 2;
    std::vector<double> A(n_A), rhs(L);
    FSIM_CUDA(cudaMemcpyAsync(A.data(), d_A, sizeof(double) * n_A, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaMemcpyAsync(rhs.data(), d_rhs, sizeof(double) * L, cudaMemcpyDeviceToHost, s->stream));
    FSIM_CUDA(cudaStreamSynchronize(s->stream));

    // the reference's solver (matrix_webgl.makeSORIterative -> fsim_jacobi_*), fp64
    fsim_jacobi *jac = nullptr;
    FSIM_TRY(fsim_jacobi_create(FSIM_SPINDLE_NPOWER, 1.0, FSIM_F64, s->device, 0, &jac));
    std::vector<double> x(L);
    double corr = 0, diff = 0;
    int32_t iters = 0;
    int rc = fsim_jacobi_set_matrix(jac, A.data());
    if (rc == FSIM_OK) rc = fsim_jacobi_set_b(jac, rhs.data());
    if (rc == FSIM_OK)
        rc = fsim_jacobi_solve(jac, FSIM_SPINDLE_TOL, FSIM_SPINDLE_SUBSTEP, FSIM_SPINDLE_MAXCHECK, &corr, &diff, &iters, x.data());
    s->launches += fsim_jacobi_launch_count(jac);
    fsim_jacobi_destroy(jac);
    FSIM_TRY(rc);
    if (!(diff <= FSIM_SPINDLE_TOL)) {
        set_error("addSpindleCuspPlasmaField: the boundary solve did not converge");
        return FSIM_ERR_RANGE;
    }

    // element strengths -> node currents (node l carries x_l - x_{l-1}), scaled by 1 - sqrt(1 - beta)
    const double scale = 1.0 - sqrt(1.0 - beta_c);
    std::vector<double> cur(L + 1), loops;
    for (int l = 0; l <= L; ++l) {
        const double plus = (l < L) ? x[l] : 0.0, minus = (l > 0) ? x[l - 1] : 0.0;
        cur[l] = scale * (plus - minus);
    }
    loops = {coil_r, 0.0, coil_I, coil_r, height, -coil_I};
    for (int l = 0; l <= L; ++l) {
        loops.insert(loops.end(), {nodes[2 * l], nodes[2 * l + 1], cur[l]});
        loops.insert(loops.end(), {nodes[2 * l], height - nodes[2 * l + 1], -cur[l]});
    }
    const int nloops = (int)(loops.size() / 3);
    FSIM_TRY(ensure_stage(s, sizeof(double) * loops.size()));
    FSIM_CUDA(cudaMemcpyAsync(s->stage, loops.data(), sizeof(double) * loops.size(), cudaMemcpyHostToDevice, s->stream));
    rc = dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        Bracket b(s, "add_loops");
        add_loops_kernel<Real><<<grid_for(s->ncell_local, 128), 128, 0, s->stream>>>(
            (Real *)s->B, s->nr, s->nz, s->row0, s->rows, (Real)(radius / (double)s->nr), (Real)(height / (double)s->nz), nloops,
            (const double *)s->stage);
        FSIM_CUDA(cudaGetLastError());
        return (int)FSIM_OK;
    });
    FSIM_TRY(rc);
    FSIM_CUDA(cudaStreamSynchronize(s->stream));  // `loops` lives on this stack frame
    s->spindle_x = x;
    s->spindle_currents = cur;
    s->spindle_A = A;
    s->spindle_rhs = rhs;
    s->spindle_iterations = iters;
    s->spindle_diff = diff;
    return FSIM_OK;
}

}  // namespace fsim
