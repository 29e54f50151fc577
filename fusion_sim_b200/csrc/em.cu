// em.cu -- EXTENSION (BASELINE.json configs[2] "Boris + Yee FDTD"; no reference
// counterpart: the reference's E and B are static, empic.js:1159-1197): electromagnetic field update on the
// reference's cylindrical (r, z) grid, driven by the current the deposit measures (sm_100a).
//
// An axisymmetric Yee mesh over the cells of the field textures: E_r on the z-faces' radial edges, E_z on the radial
// faces' axial edges, B_theta at the cell centres (the TM set); E_theta at the cell corners, B_r and B_z on the edges
// between them (the TE set).  One fsim_em_step() = one leap-frog step of dt: B from curl E, then E from curl B minus
// J/eps0 with J = q w v-moment / cell volume taken from moments 1 of the deposit (programMoments01, empic.js:1006),
// then the cell-centred E and B0 + B the push gathers, then precalc().  Perfectly conducting wall and end plates,
// regular axis.  The specification -- mesh, coefficients, operation order -- is in include/fusionsim.h
// (fsim_em_step); the test suite holds a CPU statement of it and the two agree bit for bit.  PARITY UNPINNED.
//
//   em_b_kernel     : B_r, B_t, B_z += curl E          reads 3 + reads/writes 3 planes   72 B/cell fp64  (HBM-bound)
//   em_e_kernel     : E_r, E_t, E_z += c^2 curl B - J  reads 3 + 3 moment planes, r/w 3   96 B/cell       (HBM-bound)
//   em_cells_kernel : edges -> cell centres            reads 6 planes + B0, writes E, B  120 B/cell       (HBM-bound)
// The stencils touch each value from at most 4 threads of neighbouring lanes/rows: L1/L2 serve the re-reads.
#include <math.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace fsim {

enum { EM_ER = 0, EM_EZ, EM_BT, EM_ET, EM_BR, EM_BZ };

template <typename Real>
struct EmArgs {
    Real *Er, *Ez, *Bt, *Et, *Br, *Bz;
    const Real *coef;            // [nr+1][6] = a1 a0 b1 b0 gR gZ
    const Real *mom;             // planar moments 1 (3 planes used) or nullptr
    int nr, nz, pitch;
    int64_t plane;
    Real kz, kr, cz, cr, cj, ax;
};

// one thread per lattice point (i, j), 0 <= i <= nr, 0 <= j <= nz: the components that live at that index
template <typename Real>
__global__ void __launch_bounds__(256)
em_b_kernel(const EmArgs<Real> a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int nr = a.nr, nz = a.nz, P = nr + 1;
    if (i > nr) return;
    const size_t oP = (size_t)j * P + i, oN = (size_t)j * nr + i;
    const Real et = a.Et[oP];
    if (j < nz) a.Br[oP] = a.Br[oP] + a.kz * (a.Et[oP + P] - et);
    if (i < nr) {
        const Real *k = a.coef + 6 * (size_t)i;
        if (j < nz) a.Bt[oN] = a.Bt[oN] - ((a.kz * (a.Er[oN + nr] - a.Er[oN])) - (a.kr * (a.Ez[oP + 1] - a.Ez[oP])));
        a.Bz[oN] = a.Bz[oN] - ((k[0] * a.Et[oP + 1]) - (k[1] * et));
    }
}

template <typename Real, bool J>
__global__ void __launch_bounds__(256)
em_e_kernel(const EmArgs<Real> a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int nr = a.nr, nz = a.nz, P = nr + 1;
    if (i > nr) return;
    const size_t oP = (size_t)j * P + i, oN = (size_t)j * nr + i;
    const Real *k = a.coef + 6 * (size_t)i;
    // current density at the centre of cell (ii, jj): g * moment
    auto jq = [&](int q, int jj, int ii) -> Real {
        if constexpr (!J) return (Real)0;
        else return a.coef[6 * (size_t)ii + (q == 2 ? 5 : 4)] * a.mom[(size_t)q * a.plane + (size_t)jj * a.pitch + ii];
    };
    if (i < nr && j >= 1 && j <= nz - 1)
        a.Er[oN] = a.Er[oN] + ((-(a.cz * (a.Bt[oN] - a.Bt[oN - nr]))) - a.cj * ((Real)0.5 * (jq(0, j - 1, i) + jq(0, j, i))));
    if (i >= 1 && i <= nr - 1 && j >= 1 && j <= nz - 1)
        a.Et[oP] = a.Et[oP] + (((a.cz * (a.Br[oP] - a.Br[oP - P])) - (a.cr * (a.Bz[oN] - a.Bz[oN - 1]))) -
                               a.cj * ((Real)0.25 * (((jq(1, j - 1, i - 1) + jq(1, j - 1, i)) + jq(1, j, i - 1)) + jq(1, j, i))));
    if (j < nz && i >= 1 && i <= nr - 1)
        a.Ez[oP] = a.Ez[oP] + (((k[2] * a.Bt[oN]) - (k[3] * a.Bt[oN - 1])) - a.cj * ((Real)0.5 * (jq(2, j, i - 1) + jq(2, j, i))));
    if (j < nz && i == 0) a.Ez[oP] = a.Ez[oP] + ((a.ax * a.Bt[oN]) - a.cj * jq(2, j, 0));
}

template <typename Real>
__global__ void __launch_bounds__(256)
em_cells_kernel(const EmArgs<Real> a, const Real *__restrict__ B0, Real *__restrict__ E, Real *__restrict__ B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int nr = a.nr, P = nr + 1;
    if (i >= nr) return;
    const size_t oP = (size_t)j * P + i, oN = (size_t)j * nr + i, c = oN;
    E[3 * c] = (Real)0.5 * (a.Er[oN] + a.Er[oN + nr]);
    E[3 * c + 1] = (Real)0.25 * (((a.Et[oP] + a.Et[oP + 1]) + a.Et[oP + P]) + a.Et[oP + P + 1]);
    E[3 * c + 2] = (Real)0.5 * (a.Ez[oP] + a.Ez[oP + 1]);
    B[3 * c] = B0[3 * c] + (Real)0.5 * (a.Br[oP] + a.Br[oP + 1]);
    B[3 * c + 1] = B0[3 * c + 1] + a.Bt[oN];
    B[3 * c + 2] = B0[3 * c + 2] + (Real)0.5 * (a.Bz[oN] + a.Bz[oN + nr]);
}

static size_t em_count(const fsim_sim *s, int f)
{
    const size_t nr = (size_t)s->nr, nz = (size_t)s->nz;
    switch (f) {
    case EM_ER: case EM_BZ: return (nz + 1) * nr;
    case EM_EZ: case EM_BR: return nz * (nr + 1);
    case EM_BT: return nz * nr;
    default: return (nz + 1) * (nr + 1);
    }
}

int em_field_index(const std::string &n)
{
    static const char *names[6] = {"Er", "Ez", "Bt", "Et", "Br", "Bz"};
    for (int f = 0; f < 6; ++f)
        if (n == names[f]) return f;
    return -1;
}

int64_t em_field_count(const fsim_sim *s, int f) { return (int64_t)em_count(s, f); }

void em_free(fsim_sim *s)
{
    for (int f = 0; f < 6; ++f) { if (s->em[f]) cudaFree(s->em[f]); s->em[f] = nullptr; }
    if (s->em_B0) cudaFree(s->em_B0);
    if (s->em_coef) cudaFree(s->em_coef);
    s->em_B0 = s->em_coef = nullptr;
}

// zero Yee fields; the static B present now is kept underneath (B = B0 + B_em from here on)
int em_init(fsim_sim *s)
{
    const double dr = s->spec.radius / (double)s->nr, dz = s->spec.height / (double)s->nz;
    if (!(FSIM_C_LIGHT * s->spec.dt * sqrt(1.0 / (dr * dr) + 1.0 / (dz * dz)) < 1.0)) {
        set_error(".dt <- the Yee update needs c dt sqrt(1/dr^2 + 1/dz^2) < 1");
        return FSIM_ERR_RANGE;
    }
    if (s->nz + 1 > 65535) {  // one block row per lattice row (gridDim.y)
        set_error(".nz <- the electromagnetic update takes at most 65534 rows");
        return FSIM_ERR_RANGE;
    }
    for (int f = 0; f < 6; ++f) {
        const size_t bytes = s->rs * em_count(s, f);
        if (!s->em[f]) FSIM_CUDA(cudaMalloc(&s->em[f], bytes));
        FSIM_CUDA(cudaMemsetAsync(s->em[f], 0, bytes, s->stream));
    }
    const size_t bB = s->rs * 3 * (size_t)s->ncell_local;
    if (!s->em_B0) FSIM_CUDA(cudaMalloc(&s->em_B0, bB));
    FSIM_CUDA(cudaMemcpyAsync(s->em_B0, s->B, bB, cudaMemcpyDeviceToDevice, s->stream));
    if (!s->em_coef) FSIM_CUDA(cudaMalloc(&s->em_coef, s->rs * 6 * (size_t)(s->nr + 1)));
    s->em_weight = nan("");  // coefficients not uploaded yet
    return FSIM_OK;
}

// per-column coefficients in host fp64, rounded to the handle's precision (specification: include/fusionsim.h)
static int em_upload_coef(fsim_sim *s, double macro_weight)
{
    const int nr = s->nr;
    const double dr = s->spec.radius / (double)nr, dz = s->spec.height / (double)s->nz, dt = s->spec.dt;
    const double c2 = FSIM_C_LIGHT * FSIM_C_LIGHT;
    std::vector<double> cd(6 * (size_t)(nr + 1));
    for (int i = 0; i <= nr; ++i) {
        const double rh = ((double)i + 0.5) * dr;
        cd[6 * i] = dt * ((double)i + 1.0) * dr / (rh * dr);
        cd[6 * i + 1] = dt * (double)i * dr / (rh * dr);
        cd[6 * i + 2] = i ? c2 * dt * rh / ((double)i * dr * dr) : 0.0;
        cd[6 * i + 3] = i ? c2 * dt * (((double)i - 0.5) * dr) / ((double)i * dr * dr) : 0.0;
        const double u = ((double)i + 0.5) / (double)nr;
        const double G = s->spec.particle_charge * macro_weight * 1000.0 * FSIM_C_LIGHT / (2.0 * FSIM_PI * u * s->spec.radius * dr * dz);
        cd[6 * i + 4] = G * s->spec.radius;
        cd[6 * i + 5] = G * s->spec.height;
    }
    if (s->prec == FSIM_F64) {
        FSIM_CUDA(cudaMemcpyAsync(s->em_coef, cd.data(), sizeof(double) * cd.size(), cudaMemcpyHostToDevice, s->stream));
    } else {
        std::vector<float> cf(cd.begin(), cd.end());
        FSIM_CUDA(cudaMemcpyAsync(s->em_coef, cf.data(), sizeof(float) * cf.size(), cudaMemcpyHostToDevice, s->stream));
    }
    FSIM_CUDA(cudaStreamSynchronize(s->stream));  // the host vectors go out of scope
    s->em_weight = macro_weight;
    return FSIM_OK;
}

int em_step(fsim_sim *s, double macro_weight, bool with_current)
{
    if (!(s->em_weight == macro_weight)) FSIM_TRY(em_upload_coef(s, macro_weight));
    return dispatch(s, [&](auto tag) {
        using Real = decltype(tag);
        const double dr = s->spec.radius / (double)s->nr, dz = s->spec.height / (double)s->nz, dt = s->spec.dt;
        const double c2 = FSIM_C_LIGHT * FSIM_C_LIGHT;
        EmArgs<Real> a;
        a.Er = (Real *)s->em[EM_ER]; a.Ez = (Real *)s->em[EM_EZ]; a.Bt = (Real *)s->em[EM_BT];
        a.Et = (Real *)s->em[EM_ET]; a.Br = (Real *)s->em[EM_BR]; a.Bz = (Real *)s->em[EM_BZ];
        a.coef = (const Real *)s->em_coef;
        a.mom = with_current ? (const Real *)s->mom : nullptr;
        a.nr = s->nr; a.nz = s->nz; a.pitch = s->pitch; a.plane = s->plane;
        a.kz = (Real)(dt / dz); a.kr = (Real)(dt / dr); a.cz = (Real)(c2 * dt / dz); a.cr = (Real)(c2 * dt / dr);
        a.cj = (Real)(dt / FSIM_EPS0); a.ax = (Real)(4.0 * c2 * dt / dr);
        const dim3 lattice((s->nr + 1 + 255) / 256, s->nz + 1), cells((s->nr + 255) / 256, s->nz);
        {
            Bracket b(s, "em_b");
            em_b_kernel<Real><<<lattice, 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        {
            Bracket b(s, "em_e");
            if (with_current) em_e_kernel<Real, true><<<lattice, 256, 0, s->stream>>>(a);
            else em_e_kernel<Real, false><<<lattice, 256, 0, s->stream>>>(a);
            FSIM_CUDA(cudaGetLastError());
        }
        {
            Bracket b(s, "em_cells");
            em_cells_kernel<Real><<<cells, 256, 0, s->stream>>>(a, (const Real *)s->em_B0, (Real *)s->E, (Real *)s->B);
            FSIM_CUDA(cudaGetLastError());
        }
        return (int)FSIM_OK;
    });
}

}  // namespace fsim
