// P1 element kernels on triangles / tetrahedra (mode F: DG0 lagged velocity) + symmetric Dirichlet
// elimination + fused cellwise velocity recovery.
//
// These replace the FFC-generated tabulate_tensor of the weak forms in
// /root/reference/src/modules/mupopp.py (line ranges cited per kernel).  One thread per cell:
// a 16-byte coalesced load of the cell's vertex ids, gathers of coordinates and nodal fields
// (L2-resident neighbours), closed-form exact integration of the barycentric monomials, and a
// contiguous [cell][i][j] store of the element tensor that the row-gather scatter (plan.cu) consumes.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <int D>
struct Geom {
    int32_t v[D + 1];     // vertex ids (geometry)
    int32_t dof[D + 1];   // P1 dof ids (== v unless the space is periodic)
    double g[D + 1][D];  // physical gradients of the barycentric (P1) basis
    double vol;
    double h;            // UFL CellDiameter: max vertex distance (mupopp.py:797,828)
};

template <int D>
__device__ __forceinline__ void load_geom(const mp_mesh &m, int64_t c, Geom<D> &G) {
    const int32_t *__restrict__ cells = m.d_cells;
    const double *__restrict__ coords = m.d_coords;
    double x[D + 1][D];
    if (D == 3) {
        const int4 q = *reinterpret_cast<const int4 *>(cells + 4 * c);
        G.v[0] = q.x; G.v[1] = q.y; G.v[2] = q.z; G.v[D] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i <= D; ++i) G.v[i] = cells[(D + 1) * c + i];
    }
    if (m.d_cell_dofs) {
#pragma unroll
        for (int i = 0; i <= D; ++i) G.dof[i] = m.d_cell_dofs[(D + 1) * c + i];
    } else {
#pragma unroll
        for (int i = 0; i <= D; ++i) G.dof[i] = G.v[i];
    }
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
        for (int a = 0; a < D; ++a) x[i][a] = coords[(int64_t)G.v[i] * D + a];
    double hmax2 = 0.0;
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
        for (int j = i + 1; j <= D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) { double d = x[i][a] - x[j][a]; s += d * d; }
            hmax2 = fmax(hmax2, s);
        }
    G.h = sqrt(hmax2);
    if (D == 2) {
        const double a = x[1][0] - x[0][0], b = x[2][0] - x[0][0];
        const double c2 = x[1][1] - x[0][1], d = x[2][1] - x[0][1];
        const double det = a * d - b * c2;
        const double inv = 1.0 / det;
        // Jinv = 1/det [[d,-b],[-c,a]] ; grad lambda_k = row k-1 of Jinv
        G.g[1][0] = d * inv;  G.g[1][1] = -b * inv;
        G.g[2][0] = -c2 * inv; G.g[2][1] = a * inv;
        G.vol = 0.5 * fabs(det);
    } else {
        double J[3][3];  // J[i][a] = x[a+1][i]-x[0][i]
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int a = 0; a < 3; ++a) J[i][a] = x[a + 1][i] - x[0][i];
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        const double inv = 1.0 / det;
        // Jinv[a][i] = cof[i][a]/det ; grad lambda_{a+1} = row a of Jinv
        G.g[1][0] = c00 * inv;
        G.g[2][0] = c01 * inv;
        G.g[3][0] = c02 * inv;
        G.g[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * inv;
        G.g[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv;
        G.g[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * inv;
        G.g[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv;
        G.g[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * inv;
        G.g[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv;
        G.vol = fabs(det) * (1.0 / 6.0);
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double s = 0.0;
#pragma unroll
        for (int i = 1; i <= D; ++i) s += G.g[i][a];
        G.g[0][a] = -s;
    }
}

template <int D> __device__ __forceinline__ void gather(const double *__restrict__ f, const Geom<D> &G, double (&out)[D + 1]) {
#pragma unroll
    for (int i = 0; i <= D; ++i) out[i] = f[G.dof[i]];
}
template <int D> __device__ __forceinline__ double sum(const double (&a)[D + 1]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i <= D; ++i) s += a[i];
    return s;
}

// exact integrals of barycentric monomials over a simplex, divided by |K|:  d! prod(alpha!) / (d+|alpha|)!
template <int D> struct Coef {
    static constexpr double c2 = 1.0 / ((D + 1) * (D + 2));
    static constexpr double c3 = c2 / (D + 3);
    static constexpr double c4 = c3 / (D + 4);
};

// (1/|K|) int lambda_i * a * b      (a, b P1 nodal)
template <int D>
__device__ __forceinline__ double T2(int i, const double (&a)[D + 1], const double (&b)[D + 1], double Sa, double Sb, double Sab) {
    return Coef<D>::c3 * (Sa * Sb + a[i] * Sb + Sa * b[i] + Sab + 2.0 * a[i] * b[i]);
}
// (1/|K|) int a*b*c
template <int D>
__device__ __forceinline__ double I3(const double (&a)[D + 1], const double (&b)[D + 1], const double (&c)[D + 1]) {
    double Sa = 0, Sb = 0, Sc = 0, Sab = 0, Sac = 0, Sbc = 0, Sabc = 0;
#pragma unroll
    for (int j = 0; j <= D; ++j) {
        Sa += a[j]; Sb += b[j]; Sc += c[j];
        Sab += a[j] * b[j]; Sac += a[j] * c[j]; Sbc += b[j] * c[j]; Sabc += a[j] * b[j] * c[j];
    }
    return Coef<D>::c3 * (Sa * Sb * Sc + Sa * Sbc + Sb * Sac + Sc * Sab + 2.0 * Sabc);
}
// (1/|K|) int lambda_i * a*b*c  (degree 4): brute force over the multi-index multiplicities (folded at compile time)
template <int D>
__device__ __forceinline__ void T3_all(const double (&a)[D + 1], const double (&b)[D + 1], const double (&c)[D + 1], double (&out)[D + 1]) {
#pragma unroll
    for (int i = 0; i <= D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j <= D; ++j)
#pragma unroll
            for (int k = 0; k <= D; ++k)
#pragma unroll
                for (int l = 0; l <= D; ++l) {
                    int cnt[D + 1];
#pragma unroll
                    for (int v = 0; v <= D; ++v) cnt[v] = (v == i) + (v == j) + (v == k) + (v == l);
                    int m = 1;
#pragma unroll
                    for (int v = 0; v <= D; ++v) m *= (cnt[v] <= 1 ? 1 : cnt[v] == 2 ? 2 : cnt[v] == 3 ? 6 : 24);
                    s += (double)m * a[j] * b[k] * c[l];
                }
        out[i] = Coef<D>::c4 * s;
    }
}

// reaction base rate c0*c1 (kind 0) or c0*c1^2 (kind 1): test-weighted integrals and mean, all divided by |K|
template <int D>
__device__ __forceinline__ void reaction_ints(int kind, const double (&c0)[D + 1], const double (&c1)[D + 1], double (&Ti)[D + 1], double &mean) {
    if (kind == 0) {
        double Sa = 0, Sb = 0, Sab = 0;
#pragma unroll
        for (int j = 0; j <= D; ++j) { Sa += c0[j]; Sb += c1[j]; Sab += c0[j] * c1[j]; }
#pragma unroll
        for (int i = 0; i <= D; ++i) Ti[i] = T2<D>(i, c0, c1, Sa, Sb, Sab);
        mean = Coef<D>::c2 * (Sa * Sb + Sab);
    } else {
        T3_all<D>(c0, c1, c1, Ti);
        mean = I3<D>(c0, c1, c1);
    }
}

__device__ __forceinline__ double tau_supg(int supg, double Pe, double Da, double h, double vn) {
    // mupopp.py:835-846
    if (supg == 1) return 1.0 / (4.0 / (Pe * h * h) + 2.0 * vn / h);
    if (supg == 2) return 1.0 / (4.0 / (Pe * h * h) + 2.0 * vn / h + Da);
    const double a = Pe * vn * h / 2.0;
    const double e2 = exp(2.0 * a);
    const double coth = (e2 + 1.0) / (e2 - 1.0);
    return 0.5 * h * (coth - 1.0 / a) / vn;
}

template <int D>
__device__ __forceinline__ void store_mat(const mp_mesh &m, double *__restrict__ emat, int64_t c, const double (&A)[D + 1][D + 1]) {
    if (m.d_row_pos) {   // row-ordered element storage: row i of this cell goes to its incidence slot (skipped if not assembled here)
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            const int32_t e = m.d_row_pos[c * (D + 1) + i];
            if (e < 0) continue;
            double *dst = emat + (int64_t)e * (D + 1);
            if (D == 3) {
                reinterpret_cast<double2 *>(dst)[0] = make_double2(A[i][0], A[i][1]);
                reinterpret_cast<double2 *>(dst)[1] = make_double2(A[i][2], A[i][3]);
            } else {
#pragma unroll
                for (int j = 0; j <= D; ++j) dst[j] = A[i][j];
            }
        }
        return;
    }
    double *dst = emat + c * (D + 1) * (D + 1);
    if (D == 3) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            reinterpret_cast<double2 *>(dst)[2 * i] = make_double2(A[i][0], A[i][1]);
            reinterpret_cast<double2 *>(dst)[2 * i + 1] = make_double2(A[i][2], A[i][3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i <= D; ++i)
#pragma unroll
            for (int j = 0; j <= D; ++j) dst[i * (D + 1) + j] = A[i][j];
    }
}
template <int D>
__device__ __forceinline__ void store_vec(const mp_mesh &m, double *__restrict__ evec, int64_t c, const double (&b)[D + 1]) {
    if (m.d_row_pos) {
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            const int32_t e = m.d_row_pos[c * (D + 1) + i];
            if (e >= 0) evec[e] = b[i];
        }
        return;
    }
    double *dst = evec + c * (D + 1);
    if (D == 3) {
        reinterpret_cast<double2 *>(dst)[0] = make_double2(b[0], b[1]);
        reinterpret_cast<double2 *>(dst)[1] = make_double2(b[2], b[3]);
    } else {
#pragma unroll
        for (int i = 0; i <= D; ++i) dst[i] = b[i];
    }
}

template <int D>
__device__ __forceinline__ double cell_K(const double *__restrict__ K, double Kconst, const Geom<D> &G) {
    if (!K) return Kconst;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i <= D; ++i) s += K[G.dof[i]];
    return s * (1.0 / (D + 1));
}

// ------------------------------------------------------------------ kernels

template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_mass(mp_mesh m, double scale, double *__restrict__ emat) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double A[D + 1][D + 1];
    const double w = scale * G.vol * Coef<D>::c2;
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
        for (int j = 0; j <= D; ++j) A[i][j] = (i == j) ? 2.0 * w : w;
    store_mat<D>(m, emat, c, A);
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_stiffness(mp_mesh m, const double *__restrict__ K, double Kconst, double scale,
                                                           double *__restrict__ emat) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    const double w = scale * cell_K<D>(K, Kconst, G) * G.vol;
    double A[D + 1][D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
        for (int j = 0; j <= D; ++j) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) s += G.g[i][a] * G.g[j][a];
            A[i][j] = w * s;
        }
    store_mat<D>(m, emat, c, A);
}

// c1 / c2 rows: qc*(cc-c0)*dx + dt*f21/(1-phi)*qc*dx ; qc1*(cc1-c1)*dx - dt*f31/(1-phi)*qc1*dx
// (mupopp.py:826, :936-937, :1234-1235) -> right-hand sides of the consistent-mass solves.
template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_reaction_rhs(mp_mesh m, mp_transport_params P, const double *__restrict__ c0n,
                                                              const double *__restrict__ c1n, const double *__restrict__ c2n,
                                                              double *__restrict__ eb1, double *__restrict__ eb2) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double a0[D + 1], a1[D + 1], Ti[D + 1], mean;
    gather<D>(c0n, G, a0);
    gather<D>(c1n, G, a1);
    reaction_ints<D>(P.reaction, a0, a1, Ti, mean);
    const double S1 = sum<D>(a1);
    double b[D + 1];
    const double k2 = P.dt * P.frac[1] * P.Da / (1.0 - P.phi);
#pragma unroll
    for (int i = 0; i <= D; ++i) b[i] = G.vol * (Coef<D>::c2 * (S1 + a1[i]) - k2 * Ti[i]);
    store_vec<D>(m, eb1, c, b);
    if (P.ncomp == 3 && eb2) {
        double a2[D + 1];
        gather<D>(c2n, G, a2);
        const double S2 = sum<D>(a2);
        const double k3 = P.dt * P.frac[2] * P.Da / (1.0 - P.phi);
#pragma unroll
        for (int i = 0; i <= D; ++i) b[i] = G.vol * (Coef<D>::c2 * (S2 + a2[i]) + k3 * Ti[i]);
        store_vec<D>(m, eb2, c, b);
    }
}

// the reaction load alone: elem_T[i] = int w_i R(c0, c1)  (R = c0 c1 or c0 c1^2).  The c1 and c2 rows are M c1' = M c1 - k2 T and
// M c2' = M c2 + k3 T, so ONE mass solve for d = -k2 M^-1 T gives both increments (TransportSolverF._solve_reaction_rows_shared);
// assembling T directly avoids the cancellation of (M c1 - k2 T) - M c1.
template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_reaction_load(mp_mesh m, mp_transport_params P, const double *__restrict__ c0n,
                                                               const double *__restrict__ c1n, double scale, double *__restrict__ eT) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double a0[D + 1], a1[D + 1], Ti[D + 1], mean;
    gather<D>(c0n, G, a0);
    gather<D>(c1n, G, a1);
    reaction_ints<D>(P.reaction, a0, a1, Ti, mean);
    double b[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) b[i] = scale * G.vol * Ti[i];
    store_vec<D>(m, eT, c, b);
}

// c0 row, Crank-Nicolson + SUPG with the lagged velocity (mupopp.py:823-848, :934-958, :1231-1258):
//   eta (c0-c0n) + dt [eta u.grad(cm) + grad(eta).grad(cm)/Pe] + dt R1/phi eta - beta dt f eta + tau (u.grad eta) r
//   r = c0-c0n + dt (u.grad(cm) + R1/phi) - beta dt f + (c1-c1n + dt R2/(1-phi)) [+ (c2-c2n - dt R3/(1-phi))]
template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_transport(mp_mesh m, mp_transport_params P, const double *__restrict__ ucell,
                                                           const double *__restrict__ c0n, const double *__restrict__ c1n,
                                                           const double *__restrict__ c2n, const double *__restrict__ c1new,
                                                           const double *__restrict__ c2new, const double *__restrict__ fsrc,
                                                           double *__restrict__ emat, double *__restrict__ evec) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double u[D], vn2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) { u[a] = ucell[c * D + a]; vn2 += u[a] * u[a]; }
    const double tau = tau_supg(P.supg, P.Pe, P.Da, G.h, sqrt(vn2));
    double ug[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) s += u[a] * G.g[i][a];
        ug[i] = s;
    }
    const double vol = G.vol, dt = P.dt;
    const double m1 = vol * (1.0 / (D + 1));
    if (emat) {
        double A[D + 1][D + 1];
#pragma unroll
        for (int i = 0; i <= D; ++i)
#pragma unroll
            for (int j = 0; j <= D; ++j) {
                double gg = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) gg += G.g[i][a] * G.g[j][a];
                A[i][j] = vol * Coef<D>::c2 * (i == j ? 2.0 : 1.0) + 0.5 * dt * m1 * ug[j] + 0.5 * dt / P.Pe * vol * gg +
                          tau * ug[i] * m1 + 0.5 * dt * tau * vol * ug[i] * ug[j];
            }
        store_mat<D>(m, emat, c, A);
    }
    if (evec) {
        double a0[D + 1], a1[D + 1], f[D + 1], Ti[D + 1], meanR;
        gather<D>(c0n, G, a0);
        gather<D>(c1n, G, a1);
        gather<D>(fsrc, G, f);
        reaction_ints<D>(P.reaction, a0, a1, Ti, meanR);
        const double S0 = sum<D>(a0), Sf = sum<D>(f), S1 = sum<D>(a1);
        double ugc0 = 0.0, gc0[D];
#pragma unroll
        for (int a = 0; a < D; ++a) gc0[a] = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            ugc0 += ug[i] * a0[i];
#pragma unroll
            for (int a = 0; a < D; ++a) gc0[a] += a0[i] * G.g[i][a];
        }
        const double inv = 1.0 / (D + 1);
        const double kR1 = dt * P.frac[0] * P.Da / P.phi;
        double rk = -S0 * inv + 0.5 * dt * ugc0 + kR1 * meanR - P.beta * dt * Sf * inv - S1 * inv +
                    dt * P.frac[1] * P.Da / (1.0 - P.phi) * meanR;
        double cpl;  // mean of the NEW c1 (+c2): enters through the SUPG residual only
        {
            double n1[D + 1];
            gather<D>(c1new, G, n1);
            cpl = sum<D>(n1) * inv;
        }
        if (P.ncomp == 3) {
            double a2[D + 1], n2[D + 1];
            gather<D>(c2n, G, a2);
            gather<D>(c2new, G, n2);
            rk += -sum<D>(a2) * inv - dt * P.frac[2] * P.Da / (1.0 - P.phi) * meanR;
            cpl += sum<D>(n2) * inv;
        }
        double b[D + 1];
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            double gg = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) gg += G.g[i][a] * gc0[a];
            const double gal = Coef<D>::c2 * (S0 + a0[i]) - 0.5 * dt * ugc0 * inv - kR1 * Ti[i] + P.beta * dt * Coef<D>::c2 * (Sf + f[i]);
            b[i] = vol * (gal - 0.5 * dt / P.Pe * gg - tau * ug[i] * (rk + cpl));
        }
        store_vec<D>(m, evec, c, b);
    }
}

// DarcyAdvection.advection_diffusion (mupopp.py:627-664): tau = h/(2|w|), first-order reaction Da*c^n
template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_adr_single(mp_mesh m, double Pe, double Da, double dt, const double *__restrict__ ucell,
                                                            const double *__restrict__ cn, double *__restrict__ emat,
                                                            double *__restrict__ evec) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double u[D], vn2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) { u[a] = ucell[c * D + a]; vn2 += u[a] * u[a]; }
    const double tau = G.h / (2.0 * sqrt(vn2));
    double ug[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) s += u[a] * G.g[i][a];
        ug[i] = s;
    }
    const double vol = G.vol, m1 = vol * (1.0 / (D + 1));
    if (emat) {
        double A[D + 1][D + 1];
#pragma unroll
        for (int i = 0; i <= D; ++i)
#pragma unroll
            for (int j = 0; j <= D; ++j) {
                double gg = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) gg += G.g[i][a] * G.g[j][a];
                A[i][j] = vol * Coef<D>::c2 * (i == j ? 2.0 : 1.0) + 0.5 * dt * m1 * ug[j] + 0.5 * dt / Pe * vol * gg +
                          tau * ug[i] * m1 + 0.5 * dt * tau * vol * ug[i] * ug[j];
            }
        store_mat<D>(m, emat, c, A);
    }
    if (evec) {
        double a0[D + 1];
        gather<D>(cn, G, a0);
        const double S0 = sum<D>(a0), inv = 1.0 / (D + 1);
        double ugc = 0.0, gc[D];
#pragma unroll
        for (int a = 0; a < D; ++a) gc[a] = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            ugc += ug[i] * a0[i];
#pragma unroll
            for (int a = 0; a < D; ++a) gc[a] += a0[i] * G.g[i][a];
        }
        const double rk = -S0 * inv + 0.5 * dt * ugc + dt * Da * S0 * inv;
        double b[D + 1];
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            double gg = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) gg += G.g[i][a] * gc[a];
            b[i] = vol * (Coef<D>::c2 * (S0 + a0[i]) * (1.0 - dt * Da) - 0.5 * dt * ugc * inv - 0.5 * dt / Pe * gg - tau * ug[i] * rk);
        }
        store_vec<D>(m, evec, c, b);
    }
}

template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_pressure_rhs(mp_mesh m, mp_transport_params P, const double *__restrict__ K,
                                                              const double *__restrict__ c0, double *__restrict__ evec) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double a0[D + 1];
    gather<D>(c0, G, a0);
    const double rho = P.rho0 + P.gamma * sum<D>(a0) * (1.0 / (D + 1));
    const double w = -P.sigma * G.vol * cell_K<D>(K, P.Kconst, G) / P.phi * rho;
    double b[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) s += P.zhat[a] * G.g[i][a];
        b[i] = w * s;
    }
    store_vec<D>(m, evec, c, b);
}

// evec[c][i] = |cell| * (ucell ? ucell[c][comp] : 1): load of the volume-weighted vertex average of a cellwise field
// (the nodal velocity the run scripts write to file, e.g. `velocity_out << u0`, run/CCS2D/CCS_3comp.py:265-266)
template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_cell_load(mp_mesh m, const double *__restrict__ ucell, int comp, double *__restrict__ evec) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    const double v = G.vol * (ucell ? ucell[c * D + comp] : 1.0);
    double b[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) b[i] = v;
    store_vec<D>(m, evec, c, b);
}

// u = -(K/phi)(grad p + sigma rho zhat): DarcyAdvection (sigma=-1) gives u = -k(grad p - drho zhat)/phi (mupopp.py:563)
template <int D>
__global__ void __launch_bounds__(kThreads) k_recover_velocity(mp_mesh m, mp_transport_params P, const double *__restrict__ K,
                                                               const double *__restrict__ p, const double *__restrict__ c0,
                                                               double *__restrict__ ucell) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double a0[D + 1], pp[D + 1];
    gather<D>(c0, G, a0);
    gather<D>(p, G, pp);
    const double rho = P.rho0 + P.gamma * sum<D>(a0) * (1.0 / (D + 1));
    const double kf = cell_K<D>(K, P.Kconst, G) / P.phi;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double gp = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) gp += pp[i] * G.g[i][a];
        ucell[c * D + a] = -kf * (gp + P.sigma * rho * P.zhat[a]);
    }
}


// ------------------------------------------------------------------ generic affine "reference tensor" element kernel
// A_e[i][j] = scale * sum_t geo_t(cell) * T0[t][i][j]   (FFC's tensor representation): P2 mass, P1xP2 divergence/gradient blocks,
// stiffness, with an optional exact nodal-P1 or cellwise coefficient.  Used for the mixed (mode R) operators of
// DarcyAdvection.darcy_bilinear (mupopp.py:620: u.v, K div(v) p, div(u) q) and the Darcy rows of the monolithic forms (:807, :1204).
// Reference-tensor element kernel (FFC's "tensor representation"): A_e[i][j] = f_c * sum_k W_c[k] * T0[k][i][j], W_c = geometry factors
// of the cell (x nodal coefficient values when the coefficient is a P1 field).  A CTA stages T0 ONCE in shared memory (north_star (1):
// "shared-memory staging of the reference element"), then works through tiles of kRefTile cells: one thread per cell computes the
// geometry factors into shared memory, after which all threads produce the tile's kRefTile x L entries with consecutive threads on
// consecutive entries -- conflict-free shared-memory reads of T0 and fully coalesced 8-byte stores (the old thread-per-cell kernel
// re-read T0 from global memory for every entry and wrote L x 8-byte strided stores per warp instruction).
// `pos` != NULL: row-ordered output for the streaming scatter, local row i of cell c goes to emat[pos[c*lr + i] * lc ...].
constexpr int kRefTile = 64;
constexpr int kRefThreads = 256;   // 16 cell groups of 4 cells x 16 entry lanes: the register tiling below needs exactly this block size
constexpr int kRefMaxK = 9 * 4;     // nT <= D*D = 9 geometry terms, times (D+1) nodal coefficient values

template <int D>
__device__ __forceinline__ int reftensor_geo(const Geom<D> &G, int kind, int dir, double (&geo)[D * D]) {
    double detJ = G.vol;
#pragma unroll
    for (int k = 2; k <= D; ++k) detJ *= k;
    if (kind == 0) { geo[0] = detJ; return 1; }
    if (kind == 1 || kind == 2) {
#pragma unroll
        for (int a = 0; a < D; ++a) geo[a] = detJ * G.g[a + 1][dir];
        return D;
    }
    if (kind == 3) {
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
            for (int b = 0; b < D; ++b) {
                double t = 0.0;
#pragma unroll
                for (int i = 0; i < D; ++i) t += G.g[a + 1][i] * G.g[b + 1][i];
                geo[a * D + b] = detJ * t;
            }
        return D * D;
    }
    // kind 4, SYMGRAD block (ca, cb) = (dir / D, dir % D) of  2 symgrad(u):symgrad(v) = delta_ab grad psi_i.grad psi_j + d_b psi_i d_a psi_j
    // (compaction.momentum_conservation, mupopp.py:361), contracted with the STIFF reference tensors
    const int ca = dir / D, cb = dir % D;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) {
            double t = G.g[a + 1][cb] * G.g[b + 1][ca];
            if (ca == cb) {
#pragma unroll
                for (int i = 0; i < D; ++i) t += G.g[a + 1][i] * G.g[b + 1][i];
            }
            geo[a * D + b] = detJ * t;
        }
    return D * D;
}

template <int D>
__global__ void __launch_bounds__(kRefThreads) k_reftensor(mp_mesh m, int kind, int dir, int L, int lc, int NK, const double *__restrict__ T0,
                                                        const double *__restrict__ coef, int coef_kind, double scale,
                                                        const int32_t *__restrict__ pos, double *__restrict__ emat) {
    extern __shared__ double sm_ref[];
    double *T0s = sm_ref;                               // [NK][L]; for coef_kind 1 the global layout is already [t][v][e] = [k][e]
    double *W = sm_ref + (((size_t)NK * L + 1) & ~(size_t)1);   // [NK + 1][kRefTile]: weights k of every cell of the tile, then the cell factor (16-byte aligned)
    for (int k = threadIdx.x; k < NK * L; k += blockDim.x) T0s[k] = T0[k];
    const int64_t ntile = (m.ncell + kRefTile - 1) / kRefTile;
    // register tile: thread (cg, el) owns cells 4 cg .. 4 cg + 3 of the tile and entries el, el + 16, el + 32, ...: per k it reads 4 entries
    // of T0 (consecutive lanes -> consecutive addresses) and its 4 cell weights (two 16-byte loads) for 16 FMAs -- 0.4 shared-memory
    // loads per FMA instead of 2 for one output per thread, which left the kernel shared-memory bound with a P1 coefficient (NK = 36)
    const int cg = threadIdx.x >> 4, el = threadIdx.x & 15;
    for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int64_t c0 = tile * kRefTile;
        const int nc = (int)min((int64_t)kRefTile, m.ncell - c0);
        __syncthreads();                                // T0s ready / the previous tile's W consumed
        if ((int)threadIdx.x < kRefTile) {
            const int cl = threadIdx.x;
            if (cl < nc) {
                const int64_t c = c0 + cl;
                Geom<D> G;
                load_geom<D>(m, c, G);
                double geo[D * D];
                const int nT = reftensor_geo<D>(G, kind, dir, geo);
                if (coef_kind == 1) {
                    for (int t = 0; t < nT; ++t)
#pragma unroll
                        for (int v = 0; v <= D; ++v) W[(t * (D + 1) + v) * kRefTile + cl] = geo[t] * coef[G.dof[v]];
                    W[NK * kRefTile + cl] = scale;
                } else {
                    for (int t = 0; t < nT; ++t) W[t * kRefTile + cl] = geo[t];
                    W[NK * kRefTile + cl] = coef_kind == 2 ? scale * coef[c] : scale;
                }
            } else {
                for (int k = 0; k <= NK; ++k) W[k * kRefTile + cl] = 0.0;
            }
        }
        __syncthreads();
        for (int eb = 0; eb < L; eb += 64) {
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
            int ee[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { ee[j] = eb + el + 16 * j; if (ee[j] >= L) ee[j] = L - 1; }     // clamped loads, masked stores
            for (int k = 0; k < NK; ++k) {
                const double2 w01 = *reinterpret_cast<const double2 *>(W + k * kRefTile + 4 * cg);
                const double2 w23 = *reinterpret_cast<const double2 *>(W + k * kRefTile + 4 * cg + 2);
                const double w[4] = {w01.x, w01.y, w23.x, w23.y};
                double t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] = T0s[k * L + ee[j]];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += w[i] * t[j];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cl = 4 * cg + i;
                if (cl >= nc) continue;
                const double f = W[NK * kRefTile + cl];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = eb + el + 16 * j;
                    if (e >= L) continue;
                    const double val = f * acc[i][j];
                    if (pos) {
                        const int ri = e / lc, cj = e - ri * lc;
                        const int32_t q = pos[(c0 + cl) * (int64_t)(L / lc) + ri];
                        if (q >= 0) emat[(int64_t)q * lc + cj] = val;
                    } else {
                        emat[(c0 + cl) * (int64_t)L + e] = val;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ c0 row with numerical quadrature (mode R)
// Same form as k_p1_transport, but the lagged velocity is a continuous Lagrange field (P2 in the reference's mixed space,
// run/CCS2D/CCS_3comp.py:206-212) evaluated at the quadrature points, so tau(|u|,h) is non-polynomial and the shared
// quadrature table (SURVEY.md finding 7) is part of the parity contract.  ncomp==1 / reaction==2 / supg==4 give
// DarcyAdvection.advection_diffusion (mupopp.py:650-664).
struct QuadView {
    int nq, nbv;
    const double *w, *lam, *vphi;
};

template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_transport_quad(mp_mesh m, mp_transport_params P, QuadView Q,
                                                                const int32_t *__restrict__ vdofs, const double *__restrict__ v0,
                                                                const double *__restrict__ v1, const double *__restrict__ v2,
                                                                const double *__restrict__ c0n, const double *__restrict__ c1n,
                                                                const double *__restrict__ c2n, const double *__restrict__ c1new,
                                                                const double *__restrict__ c2new, const double *__restrict__ fsrc,
                                                                double *__restrict__ emat, double *__restrict__ evec) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double detJ = G.vol;
#pragma unroll
    for (int k = 2; k <= D; ++k) detJ *= k;
    const double dt = P.dt;
    double a0[D + 1], a1[D + 1], a2[D + 1], n1[D + 1], n2[D + 1], f[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) { a0[i] = a1[i] = a2[i] = n1[i] = n2[i] = f[i] = 0.0; }
    if (evec) {
        gather<D>(c0n, G, a0);
        if (fsrc) gather<D>(fsrc, G, f);
        if (P.ncomp >= 2) { gather<D>(c1n, G, a1); gather<D>(c1new, G, n1); }
        if (P.ncomp == 3) { gather<D>(c2n, G, a2); gather<D>(c2new, G, n2); }
    } else if (P.supg == 5) {
        gather<D>(c1n, G, a1);
    }
    double gc0[D], gg[D + 1][D + 1];
#pragma unroll
    for (int a = 0; a < D; ++a) {
        gc0[a] = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) gc0[a] += a0[i] * G.g[i][a];
    }
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
        for (int j = 0; j <= D; ++j) {
            double t = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) t += G.g[i][a] * G.g[j][a];
            gg[i][j] = t;
        }
    double A[D + 1][D + 1], b[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) {
        b[i] = 0.0;
#pragma unroll
        for (int j = 0; j <= D; ++j) A[i][j] = 0.0;
    }
    const int32_t *vd = vdofs + c * (int64_t)Q.nbv;
    const double kR1 = dt * P.frac[0] * P.Da / P.phi;
    const double kR2 = dt * P.frac[1] * P.Da / (1.0 - P.phi);
    const double kR3 = dt * P.frac[2] * P.Da / (1.0 - P.phi);
    for (int q = 0; q < Q.nq; ++q) {
        const double w = Q.w[q] * detJ;
        double lam[D + 1];
#pragma unroll
        for (int i = 0; i <= D; ++i) lam[i] = Q.lam[q * (D + 1) + i];
        double u[D];
#pragma unroll
        for (int a = 0; a < D; ++a) u[a] = 0.0;
        for (int k = 0; k < Q.nbv; ++k) {
            const double ph = Q.vphi[q * Q.nbv + k];
            const int32_t dk = vd[k];
            u[0] += ph * v0[dk];
            u[1] += ph * v1[dk];
            if (D == 3) u[D - 1] += ph * v2[dk];
        }
        double vn2 = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) vn2 += u[a] * u[a];
        const double vn = sqrt(vn2);
        double tau;
        if (P.supg == 4) tau = G.h / (2.0 * vn);                       // mupopp.py:663
        else if (P.supg == 5) {                                        // mupopp.py:738: Sendur + Da*c1^n(x)/phi (np.max is a UFL no-op)
            double c1p = 0.0;
#pragma unroll
            for (int i = 0; i <= D; ++i) c1p += lam[i] * a1[i];
            tau = 1.0 / (4.0 / (P.Pe * G.h * G.h) + 2.0 * vn / G.h + P.Da * c1p / P.phi);
        } else tau = tau_supg(P.supg, P.Pe, P.Da, G.h, vn);
        double ug[D + 1];
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            double t = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) t += u[a] * G.g[i][a];
            ug[i] = t;
        }
        if (emat) {
#pragma unroll
            for (int i = 0; i <= D; ++i)
#pragma unroll
                for (int j = 0; j <= D; ++j)
                    A[i][j] += w * (lam[i] * lam[j] + 0.5 * dt * lam[i] * ug[j] + 0.5 * dt / P.Pe * gg[i][j] +
                                    tau * ug[i] * (lam[j] + 0.5 * dt * ug[j]));
        }
        if (evec) {
            double c0q = 0.0, c1q = 0.0, c2q = 0.0, n1q = 0.0, n2q = 0.0, fq = 0.0, ugc0 = 0.0;
#pragma unroll
            for (int i = 0; i <= D; ++i) {
                c0q += lam[i] * a0[i]; c1q += lam[i] * a1[i]; c2q += lam[i] * a2[i];
                n1q += lam[i] * n1[i]; n2q += lam[i] * n2[i]; fq += lam[i] * f[i];
                ugc0 += ug[i] * a0[i];
            }
            const double R = P.reaction == 2 ? P.Da * c0q : (P.reaction == 0 ? P.Da * c0q * c1q : P.Da * c0q * c1q * c1q);
            // dt*R for the single-component form (mupopp.py:650: f = Da*u0), dt*frac0*R/phi otherwise (mupopp.py:825)
            const double R1 = P.reaction == 2 ? dt * R : kR1 / P.Da * R;
            const double gal = c0q - 0.5 * dt * ugc0 - R1 + P.beta * dt * fq;
            double rk = -c0q + 0.5 * dt * ugc0 + R1 - P.beta * dt * fq;
            double cpl = 0.0;
            if (P.ncomp >= 2) { rk += -c1q + kR2 / P.Da * R; cpl += n1q; }
            if (P.ncomp == 3) { rk += -c2q - kR3 / P.Da * R; cpl += n2q; }
#pragma unroll
            for (int i = 0; i <= D; ++i) {
                double gd = 0.0;
#pragma unroll
                for (int a = 0; a < D; ++a) gd += G.g[i][a] * gc0[a];
                b[i] += w * (lam[i] * gal - 0.5 * dt / P.Pe * gd - tau * ug[i] * (rk + cpl));
            }
        }
    }
    if (emat) store_mat<D>(m, emat, c, A);
    if (evec) store_vec<D>(m, evec, c, b);
}


// ------------------------------------------------------------------ compaction forms (mupopp.py:206-382), coefficient kernels
__device__ __forceinline__ double compaction_fn(int fn, double phi, double dL) {
    if (fn == 0) return dL * phi * pow(phi, 1.0 / (phi * phi * phi));   // dL*phi*phi**m3, mupopp.py:365 (precedence quirk, ~0)
    if (fn == 1) return dL / phi;                                       // dL*phi*phi*m3, mupopp.py:380
    return 3.0 * phi / (2.0 * (1.0 - phi) * (2.0 - phi));               // eta = 1/alpha, mupopp.py:346-347
}

// out[c] = (1/|K|) int_K f(phi): cellwise coefficient of the P1 stiffness blocks (grad p.grad q is constant per cell)
template <int D>
__global__ void __launch_bounds__(kThreads) k_cell_fn(mp_mesh m, QuadView Q, int fn, double dL, const double *__restrict__ phi,
                                                      double *__restrict__ out) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    double ph[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) ph[i] = phi[(m.d_cell_dofs ? m.d_cell_dofs : m.d_cells)[(D + 1) * c + i]];
    double s = 0.0;
    for (int q = 0; q < Q.nq; ++q) {
        double pq = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) pq += Q.lam[q * (D + 1) + i] * ph[i];
        s += Q.w[q] * compaction_fn(fn, pq, dL);
    }
    double fac = 1.0;
#pragma unroll
    for (int k = 2; k <= D; ++k) fac *= k;
    out[c] = s * fac;
}

// emat[c][i][j] = scale * int f(phi) lambda_i lambda_j
template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_mass_fn(mp_mesh m, QuadView Q, int fn, double dL, double scale,
                                                         const double *__restrict__ phi, double *__restrict__ emat) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double detJ = G.vol;
#pragma unroll
    for (int k = 2; k <= D; ++k) detJ *= k;
    double ph[D + 1], A[D + 1][D + 1];
    gather<D>(phi, G, ph);
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
        for (int j = 0; j <= D; ++j) A[i][j] = 0.0;
    for (int q = 0; q < Q.nq; ++q) {
        double lam[D + 1], pq = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) { lam[i] = Q.lam[q * (D + 1) + i]; pq += lam[i] * ph[i]; }
        const double w = scale * Q.w[q] * detJ * compaction_fn(fn, pq, dL);
#pragma unroll
        for (int i = 0; i <= D; ++i)
#pragma unroll
            for (int j = 0; j <= D; ++j) A[i][j] += w * lam[i] * lam[j];
    }
    store_mat<D>(m, emat, c, A);
}

// rhs of the momentum form (mupopp.py:356-357, 374-375):  evec_u[a][c][i] = -int h_a psi_i,  evec_p[c][i] = -int da R buoy gam/(1 - R buoy) lambda_i
//   h = dL (1-phi) (chi(phi) grad(phi)/B - R buoy zhat) + da grad(4 gam / (3 phi))
struct CompactionPrm { double da, R, B, costerm, dL, zhat[3]; };

template <int D>
__global__ void __launch_bounds__(kThreads) k_compaction_rhs(mp_mesh m, QuadView Q, CompactionPrm P, const double *__restrict__ phi,
                                                             const double *__restrict__ gam, const double *__restrict__ buoy,
                                                             double *__restrict__ evu0, double *__restrict__ evu1, double *__restrict__ evu2,
                                                             double *__restrict__ evp) {
    constexpr int NB2 = D == 2 ? 6 : 10;
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double detJ = G.vol;
#pragma unroll
    for (int k = 2; k <= D; ++k) detJ *= k;
    double ph[D + 1], ga[D + 1], bu[D + 1], gphi[D], ggam[D];
    gather<D>(phi, G, ph);
    gather<D>(gam, G, ga);
    gather<D>(buoy, G, bu);
#pragma unroll
    for (int a = 0; a < D; ++a) {
        gphi[a] = 0.0; ggam[a] = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) { gphi[a] += ph[i] * G.g[i][a]; ggam[a] += ga[i] * G.g[i][a]; }
    }
    double bvec[D][NB2], bp[D + 1];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int i = 0; i < NB2; ++i) bvec[a][i] = 0.0;
#pragma unroll
    for (int i = 0; i <= D; ++i) bp[i] = 0.0;
    for (int q = 0; q < Q.nq; ++q) {
        double lam[D + 1], pq = 0.0, gq = 0.0, bq = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) { lam[i] = Q.lam[q * (D + 1) + i]; pq += lam[i] * ph[i]; gq += lam[i] * ga[i]; bq += lam[i] * bu[i]; }
        const double w = Q.w[q] * detJ;
        const double chi = P.costerm * (2.0 * 249.0 + 20.0 * (-8065.0) * pq * pq * pq + 12.0 * 6149.0 * pq * pq + 6.0 * (-1778.0) * pq);
        double h[D];
#pragma unroll
        for (int a = 0; a < D; ++a)
            h[a] = P.dL * (1.0 - pq) * (chi * gphi[a] / P.B - P.R * bq * P.zhat[a]) + P.da * (4.0 / 3.0) * (ggam[a] / pq - gq * gphi[a] / (pq * pq));
#pragma unroll
        for (int i = 0; i < NB2; ++i) {
            const double psi = Q.vphi[q * NB2 + i];
#pragma unroll
            for (int a = 0; a < D; ++a) bvec[a][i] -= w * h[a] * psi;
        }
        const double fq = P.da * P.R * bq * gq / (1.0 - P.R * bq);
#pragma unroll
        for (int i = 0; i <= D; ++i) bp[i] -= w * fq * lam[i];
    }
    double *dst[3] = {evu0, evu1, evu2};
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int i = 0; i < NB2; ++i) dst[a][c * NB2 + i] = bvec[a][i];
    store_vec<D>(m, evp, c, bp);
}

// porosity row (compaction.mass_conservation, mupopp.py:217-259) with a P1 porosity and a P2 velocity:
//   w (phi1 - phi0 + dt (u.grad(phi_mid) - (1 - phi_mid) div u) - dt da gam) + (keff/|u|^2) (u.grad w) (phi1 - phi0 + dt (u.grad(phi_mid) - da gam))
template <int D>
__global__ void __launch_bounds__(kThreads) k_p1_porosity_quad(mp_mesh m, QuadView Q, const double *__restrict__ vdphi, double da, double dt,
                                                               const int32_t *__restrict__ vdofs, const double *__restrict__ v0,
                                                               const double *__restrict__ v1, const double *__restrict__ v2,
                                                               const double *__restrict__ phi0, const double *__restrict__ gam,
                                                               double *__restrict__ emat, double *__restrict__ evec) {
    constexpr int NB2 = D == 2 ? 6 : 10;
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= m.ncell) return;
    Geom<D> G;
    load_geom<D>(m, c, G);
    double detJ = G.vol;
#pragma unroll
    for (int k = 2; k <= D; ++k) detJ *= k;
    double p0[D + 1], ga[D + 1];
    gather<D>(phi0, G, p0);
    gather<D>(gam, G, ga);
    double uc[D][NB2];
    const int32_t *vd = vdofs + c * (int64_t)NB2;
#pragma unroll
    for (int k = 0; k < NB2; ++k) {
        const int32_t dk = vd[k];
        uc[0][k] = v0[dk];
        uc[1][k] = v1[dk];
        if (D == 3) uc[D - 1][k] = v2[dk];
    }
    double A[D + 1][D + 1], b[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) {
        b[i] = 0.0;
#pragma unroll
        for (int j = 0; j <= D; ++j) A[i][j] = 0.0;
    }
    for (int q = 0; q < Q.nq; ++q) {
        const double w = Q.w[q] * detJ;
        double lam[D + 1], u[D], divu = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) lam[i] = Q.lam[q * (D + 1) + i];
#pragma unroll
        for (int a = 0; a < D; ++a) u[a] = 0.0;
        for (int k = 0; k < NB2; ++k) {
            const double ph = Q.vphi[q * NB2 + k];
            double gk[D];   // physical gradient of the P2 basis function k: sum_alpha dpsi/dxi_alpha * Jinv[alpha][.]
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double t = 0.0;
#pragma unroll
                for (int al = 0; al < D; ++al) t += vdphi[(q * NB2 + k) * D + al] * G.g[al + 1][a];
                gk[a] = t;
            }
#pragma unroll
            for (int a = 0; a < D; ++a) { u[a] += ph * uc[a][k]; divu += gk[a] * uc[a][k]; }
        }
        double un2 = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) un2 += u[a] * u[a];
        const double keff = fmax(0.5 * G.h * sqrt(un2) - 1.0, 0.0);
        const double st = keff > 0.0 ? keff / un2 : 0.0;
        double ug[D + 1], ph0 = 0.0, gq = 0.0, ugp0 = 0.0;
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            double t = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) t += u[a] * G.g[i][a];
            ug[i] = t;
            ph0 += lam[i] * p0[i]; gq += lam[i] * ga[i]; ugp0 += t * p0[i];
        }
        if (emat) {
#pragma unroll
            for (int i = 0; i <= D; ++i)
#pragma unroll
                for (int j = 0; j <= D; ++j)
                    A[i][j] += w * (lam[i] * lam[j] + 0.5 * dt * lam[i] * ug[j] + 0.5 * dt * divu * lam[i] * lam[j] +
                                    st * ug[i] * (lam[j] + 0.5 * dt * ug[j]));
        }
        if (evec) {
            const double gal = ph0 - 0.5 * dt * ugp0 + dt * divu - 0.5 * dt * ph0 * divu + dt * da * gq;
            const double rk = -ph0 + 0.5 * dt * ugp0 - dt * da * gq;
#pragma unroll
            for (int i = 0; i <= D; ++i) b[i] += w * (lam[i] * gal - st * ug[i] * rk);
        }
    }
    if (emat) store_mat<D>(m, emat, c, A);
    if (evec) store_vec<D>(m, evec, c, b);
}

// porosity row with a P2 porosity space (X = CG2: benchmark/soliton/soliton.py:120, benchmark/sphere_3D_pure_shear/sphere.py:76), P2 velocity.
// Same form as above plus the second-derivative term of the SUPG residual, -div(grad(phi_mid)) (mupopp.py:250-251), which does not
// vanish for P2: the Laplacian of a P2 basis function is constant on an affine cell (vertex i: 4 |grad lam_i|^2, edge (a,b):
// 8 grad lam_a . grad lam_b).  ONE WARP PER CELL: the element matrix has up to 10 x 10 entries, far too many accumulators for a
// thread per cell; lane e owns entries e, e+32, ... and the per-point basis values live in the warp's shared-memory strip.
template <int D>
__global__ void __launch_bounds__(256) k_p2_porosity_quad(mp_mesh m, QuadView Q, const double *__restrict__ vdphi, double da, double dt,
                                                          const int32_t *__restrict__ dofs2, const double *__restrict__ v0,
                                                          const double *__restrict__ v1, const double *__restrict__ v2,
                                                          const double *__restrict__ phi0, const double *__restrict__ gam,
                                                          double *__restrict__ emat, double *__restrict__ evec) {
    constexpr int NB = D == 2 ? 6 : 10;
    constexpr int NE = NB * NB;
    constexpr int EPL = (NE + 31) / 32;                 // matrix entries per lane
    __shared__ double s_psi[8][NB], s_ug[8][NB], s_lap[8][NB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (c >= m.ncell) return;                            // warp-uniform
    Geom<D> G;
    load_geom<D>(m, c, G);
    double detJ = G.vol;
#pragma unroll
    for (int k = 2; k <= D; ++k) detJ *= k;
    // local coefficients of basis function `lane` (lanes >= NB carry zeros)
    const bool has = lane < NB;
    const int32_t dof = has ? dofs2[c * (int64_t)NB + lane] : 0;
    const double ub[3] = {has ? v0[dof] : 0.0, has ? v1[dof] : 0.0, (has && D == 3) ? v2[dof] : 0.0};
    const double p0b = has ? phi0[dof] : 0.0, gab = has ? gam[dof] : 0.0;
    // Laplacian of the P2 basis (constant per cell)
    if (has) {
        double l;
        if (lane <= D) {
            l = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) l += G.g[lane][a] * G.g[lane][a];
            l *= 4.0;
        } else {
            int ea, eb;
            const int e = lane - (D + 1);
            if (D == 2) { ea = e == 0 ? 1 : 0; eb = e == 2 ? 1 : 2; }                     // UFC triangle edges (1,2) (0,2) (0,1)
            else { const int ta[6] = {2, 1, 1, 0, 0, 0}, tb[6] = {3, 3, 2, 3, 2, 1}; ea = ta[e]; eb = tb[e]; }   // UFC tetrahedron edges
            l = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) l += G.g[ea][a] * G.g[eb][a];
            l *= 8.0;
        }
        s_lap[wib][lane] = l;
    }
    __syncwarp();
    double lap0 = has ? s_lap[wib][lane] * p0b : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lap0 += __shfl_xor_sync(0xffffffffu, lap0, o);
    double A[EPL], bi = 0.0;
#pragma unroll
    for (int k = 0; k < EPL; ++k) A[k] = 0.0;
    for (int q = 0; q < Q.nq; ++q) {
        const double w = Q.w[q] * detJ;
        double psi = 0.0, gk[D];
#pragma unroll
        for (int a = 0; a < D; ++a) gk[a] = 0.0;
        if (has) {
            psi = Q.vphi[q * NB + lane];
#pragma unroll
            for (int a = 0; a < D; ++a) {
                double t = 0.0;
#pragma unroll
                for (int al = 0; al < D; ++al) t += vdphi[(q * NB + lane) * D + al] * G.g[al + 1][a];
                gk[a] = t;
            }
        }
        // u, div u, phi0, gam at the point: warp sums over the basis
        double r[D + 3];
#pragma unroll
        for (int a = 0; a < D; ++a) r[a] = psi * ub[a];
        r[D] = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) r[D] += gk[a] * ub[a];
        r[D + 1] = psi * p0b;
        r[D + 2] = psi * gab;
#pragma unroll
        for (int k = 0; k < D + 3; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r[k] += __shfl_xor_sync(0xffffffffu, r[k], o);
        const double divu = r[D], ph0 = r[D + 1], gq = r[D + 2];
        double un2 = 0.0, ug = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) { un2 += r[a] * r[a]; ug += r[a] * gk[a]; }
        const double keff = fmax(0.5 * G.h * sqrt(un2) - 1.0, 0.0);
        const double st = keff > 0.0 ? keff / un2 : 0.0;
        double ugp0 = ug * p0b;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ugp0 += __shfl_xor_sync(0xffffffffu, ugp0, o);
        __syncwarp();
        if (has) { s_psi[wib][lane] = psi; s_ug[wib][lane] = ug; }
        __syncwarp();
        if (emat) {
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const int e = lane + 32 * k;
                if (e < NE) {
                    const int i = e / NB, j = e % NB;
                    const double pi = s_psi[wib][i], pj = s_psi[wib][j], ui = s_ug[wib][i], uj = s_ug[wib][j];
                    A[k] += w * (pi * pj * (1.0 + 0.5 * dt * divu) + 0.5 * dt * pi * uj + st * ui * (pj + 0.5 * dt * (uj - s_lap[wib][j])));
                }
            }
        }
        if (evec && has) {
            const double gal = ph0 - 0.5 * dt * ugp0 + dt * divu - 0.5 * dt * ph0 * divu + dt * da * gq;
            const double rk = -ph0 + 0.5 * dt * (ugp0 - lap0) - dt * da * gq;
            bi += w * (psi * gal - st * ug * rk);
        }
    }
    if (emat) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int e = lane + 32 * k;
            if (e < NE) emat[c * (int64_t)NE + e] = A[k];
        }
    }
    if (evec && has) evec[c * (int64_t)NB + lane] = bi;
}

// ------------------------------------------------------------------ fused ROW-WISE assembly of the c0 row
// One thread per CSR row.  The thread walks the row's incident cells in ascending cell id (the plan's incidence list -- the same
// summation order as the scatter, so the result is deterministic), recomputes the cell's contribution to ITS row only (one row of
// the element matrix of k_p1_transport + one entry of the element vector), accumulates the row in a shared-memory strip
// ([slot][thread]: bank-conflict free) and then writes CSR values, SELL mirror, right-hand side with the symmetric Dirichlet
// elimination, Jacobi diagonal and line-preconditioner bands in one pass.  No element tensor is stored: the 6.1 GB write + 6.1 GB
// read of `emat` at 200^3 (and the scatter / Dirichlet / Jacobi / SELL-fill / band-extract passes over the matrix) disappear; each
// cell's geometry is recomputed (d+1) times instead, from cache (neighbouring rows share their cells).
template <int D>
__device__ __forceinline__ double pick(const double (&a)[D + 1], int i) {
    double v = a[0];
#pragma unroll
    for (int k = 1; k <= D; ++k) v = (i == k) ? a[k] : v;
    return v;
}

struct RowAsmOut {
    const int32_t *rowptr, *inc_ptr, *inc, *sell_ptr;
    const uint16_t *slot;
    const uint8_t *isbc;
    const double *g;
    double *vals, *sval, *rhs, *dinv, *lo, *diag, *up;
    int32_t nrow, stride;
};

template <int D>
__global__ void __launch_bounds__(128) k_p1_transport_rows(mp_mesh m, mp_transport_params P, RowAsmOut O, const double *__restrict__ ucell,
                                                           const double *__restrict__ c0n, const double *__restrict__ c1n,
                                                           const double *__restrict__ c2n, const double *__restrict__ c1new,
                                                           const double *__restrict__ c2new, const double *__restrict__ fsrc) {
    extern __shared__ double strip[];                    // [max_row_nnz][blockDim.x]
    const int T = blockDim.x;
    const int32_t r = blockIdx.x * T + threadIdx.x;
    const bool active = r < O.nrow;
    const int32_t rb = active ? O.rowptr[r] : 0, len = active ? O.rowptr[r + 1] - rb : 0;
    double *acc = strip + threadIdx.x;
    for (int k = 0; k < len; ++k) acc[k * T] = 0.0;
    const bool rbc = active && O.isbc[r] != 0;
    double bsum = 0.0, elim = 0.0;
    int sdiag = -1, slo = -1, sup = -1;
    const int32_t e0 = active ? O.inc_ptr[r] : 0, e1 = active ? O.inc_ptr[r + 1] : 0;
    const double dt = P.dt, inv = 1.0 / (D + 1);
    for (int32_t e = e0; e < e1; ++e) {
        const int32_t id = O.inc[e];
        const int64_t c = id / (D + 1);
        const int li = id - (int32_t)c * (D + 1);
        Geom<D> G;
        load_geom<D>(m, c, G);
        double u[D], vn2 = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) { u[a] = ucell[c * D + a]; vn2 += u[a] * u[a]; }
        const double tau = tau_supg(P.supg, P.Pe, P.Da, G.h, sqrt(vn2));
        double ug[D + 1], gi[D];
#pragma unroll
        for (int i = 0; i <= D; ++i) {
            double t = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) t += u[a] * G.g[i][a];
            ug[i] = t;
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double col[D + 1];
#pragma unroll
            for (int i = 0; i <= D; ++i) col[i] = G.g[i][a];
            gi[a] = pick<D>(col, li);
        }
        const double ugi = pick<D>(ug, li);
        const double vol = G.vol, m1 = vol * inv;
        uint16_t sl[D + 1];
        if (D == 3) {
            const ushort4 q = *reinterpret_cast<const ushort4 *>(O.slot + (int64_t)e * 4);
            sl[0] = q.x; sl[1] = q.y; sl[2] = q.z; sl[D] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j <= D; ++j) sl[j] = O.slot[(int64_t)e * (D + 1) + j];
        }
        // ---- matrix row (k_p1_transport, row li)
#pragma unroll
        for (int j = 0; j <= D; ++j) {
            double gg = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) gg += gi[a] * G.g[j][a];
            const double aij = vol * Coef<D>::c2 * (li == j ? 2.0 : 1.0) + 0.5 * dt * m1 * ug[j] + 0.5 * dt / P.Pe * vol * gg + tau * ugi * m1 +
                               0.5 * dt * tau * vol * ugi * ug[j];
            const int32_t dj = G.dof[j];
            if (j == li) sdiag = sl[j];
            if (O.stride > 0) {
                if (dj == r - O.stride) slo = sl[j];
                if (dj == r + O.stride && dj < O.nrow) sup = sl[j];
            }
            if (sl[j] == 0xffff) continue;
            if (rbc) continue;
            if (dj != r && O.isbc[dj]) elim += aij * O.g[dj];
            else acc[(int)sl[j] * T] += aij;
        }
        // ---- rhs entry (k_p1_transport, entry li)
        {
            double a0[D + 1], a1[D + 1], f[D + 1], Ti[D + 1], meanR;
            gather<D>(c0n, G, a0);
            gather<D>(c1n, G, a1);
            gather<D>(fsrc, G, f);
            reaction_ints<D>(P.reaction, a0, a1, Ti, meanR);
            const double S0 = sum<D>(a0), Sf = sum<D>(f), S1 = sum<D>(a1);
            double ugc0 = 0.0, gc0[D];
#pragma unroll
            for (int a = 0; a < D; ++a) gc0[a] = 0.0;
#pragma unroll
            for (int i = 0; i <= D; ++i) {
                ugc0 += ug[i] * a0[i];
#pragma unroll
                for (int a = 0; a < D; ++a) gc0[a] += a0[i] * G.g[i][a];
            }
            const double kR1 = dt * P.frac[0] * P.Da / P.phi;
            double rk = -S0 * inv + 0.5 * dt * ugc0 + kR1 * meanR - P.beta * dt * Sf * inv - S1 * inv + dt * P.frac[1] * P.Da / (1.0 - P.phi) * meanR;
            double cpl;
            {
                double n1[D + 1];
                gather<D>(c1new, G, n1);
                cpl = sum<D>(n1) * inv;
            }
            if (P.ncomp == 3) {
                double a2[D + 1], n2[D + 1];
                gather<D>(c2n, G, a2);
                gather<D>(c2new, G, n2);
                rk += -sum<D>(a2) * inv - dt * P.frac[2] * P.Da / (1.0 - P.phi) * meanR;
                cpl += sum<D>(n2) * inv;
            }
            double gg = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) gg += gi[a] * gc0[a];
            const double gal = Coef<D>::c2 * (S0 + pick<D>(a0, li)) - 0.5 * dt * ugc0 * inv - kR1 * pick<D>(Ti, li) +
                               P.beta * dt * Coef<D>::c2 * (Sf + pick<D>(f, li));
            bsum += vol * (gal - 0.5 * dt / P.Pe * gg - tau * ugi * (rk + cpl));
        }
    }
    if (active) {
        if (rbc && sdiag >= 0 && sdiag != 0xffff) acc[sdiag * T] = 1.0;
        const double d = (sdiag >= 0 && sdiag != 0xffff) ? acc[sdiag * T] : 0.0;
        O.rhs[r] = rbc ? O.g[r] : bsum - elim;
        if (O.dinv) O.dinv[r] = d != 0.0 ? 1.0 / d : 1.0;
        if (O.stride > 0) {
            O.diag[r] = d;
            O.lo[r] = (slo >= 0 && slo != 0xffff) ? acc[slo * T] : 0.0;
            O.up[r] = (sup >= 0 && sup != 0xffff) ? acc[sup * T] : 0.0;
        }
        for (int k = 0; k < len; ++k) O.vals[rb + k] = acc[k * T];
    }
    if (O.sval) {          // SELL-32: the warp is the slice, entries [j][lane]: coalesced 256-byte rows
        const int lane = threadIdx.x & 31;
        const int32_t s = (blockIdx.x * T + threadIdx.x) >> 5;
        if ((int64_t)s * 32 < O.nrow) {
            const int32_t b = O.sell_ptr[s], w = (O.sell_ptr[s + 1] - b) >> 5;
            for (int j = 0; j < w; ++j) O.sval[b + 32 * j + lane] = j < len ? acc[j * T] : 0.0;
        }
    }
}

// Symmetric Dirichlet elimination, 8 lanes per row, fixed reduction order.
__global__ void __launch_bounds__(256) k_dirichlet(int64_t nrow, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                   double *__restrict__ vals, double *__restrict__ rhs,
                                                   const uint8_t *__restrict__ isbc, const double *__restrict__ g) {
    constexpr int S = 8;
    const int lane = threadIdx.x & 31, sl = lane % S;
    int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / S;
    const bool active = row < nrow;
    double acc = 0.0;
    bool rbc = false;
    if (active && !vals) rbc = isbc[row] != 0;  // rhs-only mode (matrix eliminated earlier, homogeneous columns)
    if (active && vals) {
        rbc = isbc[row] != 0;
        const int32_t b = rowptr[row], e = rowptr[row + 1];
        for (int32_t k = b + sl; k < e; k += S) {
            const int32_t j = colidx[k];
            if (rbc) {
                vals[k] = (j == row) ? 1.0 : 0.0;
            } else if (isbc[j]) {
                acc += vals[k] * g[j];
                vals[k] = 0.0;
            }
        }
    }
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (active && sl == 0 && rhs) rhs[row] = rbc ? g[row] : rhs[row] - acc;
}

#define DISPATCH_DIM(mesh, KERNEL, ...)                                                                        \
    do {                                                                                                       \
        unsigned grid = (unsigned)cdiv((mesh)->ncell, kThreads);                                               \
        if (grid == 0) return 0;                                                                               \
        if ((mesh)->dim == 2) KERNEL<2><<<grid, kThreads, 0, ctx->stream>>>(*(mesh), __VA_ARGS__);             \
        else KERNEL<3><<<grid, kThreads, 0, ctx->stream>>>(*(mesh), __VA_ARGS__);                              \
        mp::count_launch(ctx);                                                                                 \
        MP_CUDA(cudaGetLastError());                                                                           \
    } while (0)

int check_mesh(const mp_ctx *ctx, const mp_mesh *m) {
    MP_CHECK(ctx && m, "null ctx/mesh");
    MP_CHECK(m->dim == 2 || m->dim == 3, "mesh dim must be 2 or 3 (got %d)", m->dim);
    MP_CHECK(m->d_cells && m->d_coords, "mesh device pointers are null");
    return 0;
}

}  // namespace

extern "C" int mp_elem_p1_mass(mp_ctx *ctx, const mp_mesh *mesh, double scale, double *d_elem_mat) {
    MP_TRY(check_mesh(ctx, mesh));
    DISPATCH_DIM(mesh, k_p1_mass, scale, d_elem_mat);
    return 0;
}

extern "C" int mp_elem_p1_stiffness(mp_ctx *ctx, const mp_mesh *mesh, const double *d_K, double Kconst, double scale, double *d_elem_mat) {
    MP_TRY(check_mesh(ctx, mesh));
    DISPATCH_DIM(mesh, k_p1_stiffness, d_K, Kconst, scale, d_elem_mat);
    return 0;
}

extern "C" int mp_elem_p1_reaction_rhs(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_c0n,
                                       const double *d_c1n, const double *d_c2n, double *d_elem_b1, double *d_elem_b2) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(prm && d_c0n && d_c1n && d_elem_b1, "mp_elem_p1_reaction_rhs: null argument");
    MP_CHECK(prm->ncomp == 2 || (d_c2n && d_elem_b2), "mp_elem_p1_reaction_rhs: ncomp==3 needs c2n and elem_b2");
    DISPATCH_DIM(mesh, k_p1_reaction_rhs, *prm, d_c0n, d_c1n, d_c2n, d_elem_b1, d_elem_b2);
    return 0;
}

extern "C" int mp_elem_p1_reaction_load(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_c0n,
                                        const double *d_c1n, double scale, double *d_elem_T) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(prm && d_c0n && d_c1n && d_elem_T, "mp_elem_p1_reaction_load: null argument");
    DISPATCH_DIM(mesh, k_p1_reaction_load, *prm, d_c0n, d_c1n, scale, d_elem_T);
    return 0;
}

extern "C" int mp_elem_p1_transport(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_ucell,
                                    const double *d_c0n, const double *d_c1n, const double *d_c2n, const double *d_c1new,
                                    const double *d_c2new, const double *d_fsrc, double *d_elem_mat, double *d_elem_vec) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(prm && d_ucell, "mp_elem_p1_transport: null argument");
    MP_CHECK(!d_elem_vec || (d_c0n && d_c1n && d_c1new && d_fsrc), "mp_elem_p1_transport: rhs needs c0n,c1n,c1new,fsrc");
    MP_CHECK(!d_elem_vec || prm->ncomp == 2 || (d_c2n && d_c2new), "mp_elem_p1_transport: ncomp==3 needs c2n,c2new");
    DISPATCH_DIM(mesh, k_p1_transport, *prm, d_ucell, d_c0n, d_c1n, d_c2n, d_c1new, d_c2new, d_fsrc, d_elem_mat, d_elem_vec);
    return 0;
}

extern "C" int mp_assemble_p1_transport_rows(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const mp_plan *plan,
                                             const double *d_ucell, const double *d_c0n, const double *d_c1n, const double *d_c2n,
                                             const double *d_c1new, const double *d_c2new, const double *d_fsrc, const uint8_t *d_isbc,
                                             const double *d_g, double *d_vals, double *d_sell_vals, double *d_rhs, double *d_dinv,
                                             int64_t line_stride, double *d_lo, double *d_diag, double *d_up) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(prm && plan && d_ucell && d_c0n && d_c1n && d_c1new && d_fsrc && d_isbc && d_g && d_vals && d_rhs,
             "mp_assemble_p1_transport_rows: null argument");
    MP_CHECK(prm->ncomp == 2 || (d_c2n && d_c2new), "mp_assemble_p1_transport_rows: ncomp==3 needs c2n,c2new");
    MP_CHECK(line_stride <= 0 || (d_lo && d_diag && d_up), "mp_assemble_p1_transport_rows: line bands requested without arrays");
    int32_t lr = 0, lc = 0, mrn = 0;
    RowAsmOut O{};
    int64_t nrow = 0;
    MP_TRY(mp_plan_incidence(plan, &nrow, &lr, &lc, &mrn, &O.rowptr, &O.inc_ptr, &O.inc, &O.slot));
    MP_CHECK(lr == mesh->dim + 1 && lc == mesh->dim + 1 && O.slot, "mp_assemble_p1_transport_rows: the plan must be a P1 x P1 matrix plan");
    int64_t nslice = 0, nent = 0;
    const int32_t *scol = nullptr;
    MP_TRY(mp_plan_sell_info(plan, &nslice, &nent, &O.sell_ptr, &scol));
    MP_CHECK(nrow < (int64_t)1 << 31 && line_stride < (int64_t)1 << 31, "mp_assemble_p1_transport_rows: 32-bit row indexing");
    O.isbc = d_isbc; O.g = d_g; O.vals = d_vals; O.sval = d_sell_vals; O.rhs = d_rhs; O.dinv = d_dinv;
    O.lo = d_lo; O.diag = d_diag; O.up = d_up; O.nrow = (int32_t)nrow; O.stride = line_stride > 0 ? (int32_t)line_stride : 0;
    if (nrow == 0) return 0;
    constexpr int T = 128;
    const size_t smem = sizeof(double) * (size_t)T * (size_t)(mrn > 0 ? mrn : 1);
    MP_CHECK(smem <= 200 * 1024, "mp_assemble_p1_transport_rows: rows of %d entries do not fit the shared-memory strip", mrn);
    const unsigned grid = (unsigned)cdiv(nrow, T);
    if (mesh->dim == 2) {
        MP_CUDA(cudaFuncSetAttribute(k_p1_transport_rows<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_p1_transport_rows<2><<<grid, T, smem, ctx->stream>>>(*mesh, *prm, O, d_ucell, d_c0n, d_c1n, d_c2n, d_c1new, d_c2new, d_fsrc);
    } else {
        MP_CUDA(cudaFuncSetAttribute(k_p1_transport_rows<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_p1_transport_rows<3><<<grid, T, smem, ctx->stream>>>(*mesh, *prm, O, d_ucell, d_c0n, d_c1n, d_c2n, d_c1new, d_c2new, d_fsrc);
    }
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mp_elem_p1_adr_single(mp_ctx *ctx, const mp_mesh *mesh, double Pe, double Da, double dt, const double *d_ucell,
                                     const double *d_cn, double *d_elem_mat, double *d_elem_vec) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(d_ucell && (!d_elem_vec || d_cn), "mp_elem_p1_adr_single: null argument");
    DISPATCH_DIM(mesh, k_p1_adr_single, Pe, Da, dt, d_ucell, d_cn, d_elem_mat, d_elem_vec);
    return 0;
}

extern "C" int mp_elem_p1_pressure_rhs(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_K,
                                       const double *d_c0, double *d_elem_vec) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(prm && d_c0 && d_elem_vec, "mp_elem_p1_pressure_rhs: null argument");
    DISPATCH_DIM(mesh, k_p1_pressure_rhs, *prm, d_K, d_c0, d_elem_vec);
    return 0;
}

extern "C" int mp_elem_p1_cell_load(mp_ctx *ctx, const mp_mesh *mesh, const double *d_ucell, int comp, double *d_elem_vec) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(d_elem_vec && comp >= 0 && comp < mesh->dim, "mp_elem_p1_cell_load: bad argument");
    DISPATCH_DIM(mesh, k_p1_cell_load, d_ucell, comp, d_elem_vec);
    return 0;
}

extern "C" int mp_recover_velocity(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_K,
                                   const double *d_p, const double *d_c0, double *d_ucell) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(prm && d_p && d_c0 && d_ucell, "mp_recover_velocity: null argument");
    DISPATCH_DIM(mesh, k_recover_velocity, *prm, d_K, d_p, d_c0, d_ucell);
    return 0;
}

extern "C" int mp_apply_dirichlet_sym(mp_ctx *ctx, int64_t nrow, const int32_t *d_rowptr, const int32_t *d_colidx, double *d_vals,
                                      double *d_rhs, const uint8_t *d_isbc, const double *d_g) {
    MP_CHECK(ctx && d_rowptr && d_colidx && d_isbc && d_g, "mp_apply_dirichlet_sym: null argument");
    MP_CHECK(d_vals || d_rhs, "mp_apply_dirichlet_sym: nothing to do (vals and rhs both null)");
    if (nrow == 0) return 0;
    k_dirichlet<<<(unsigned)cdiv(nrow * 8, 256), 256, 0, ctx->stream>>>(nrow, d_rowptr, d_colidx, d_vals, d_rhs, d_isbc, d_g);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

static int reftensor_impl(mp_ctx *ctx, const mp_mesh *mesh, int kind, int dir, int lr, int lc, const double *d_T0, const double *d_coef,
                          int coef_kind, double scale, const int32_t *d_row_pos, double *d_elem_mat) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(kind >= 0 && kind <= 4 && dir >= 0 && dir < (kind == 4 ? mesh->dim * mesh->dim : mesh->dim) && lr > 0 && lc > 0 && d_T0 && d_elem_mat,
             "mp_elem_reftensor: bad argument");
    MP_CHECK(coef_kind == 0 || d_coef, "mp_elem_reftensor: coefficient pointer is null");
    if (mesh->ncell == 0) return 0;
    const int D = mesh->dim, L = lr * lc;
    const int nT = kind == 0 ? 1 : (kind <= 2 ? D : D * D);
    const int NK = coef_kind == 1 ? nT * (D + 1) : nT;
    const size_t smem = sizeof(double) * ((size_t)NK * L + (size_t)kRefTile * (NK + 1) + 2);
    MP_CHECK(smem <= 220 * 1024, "mp_elem_reftensor: a %d x %d block with %d reference tensors does not fit in shared memory", lr, lc, NK);
    const int64_t ntile = cdiv(mesh->ncell, kRefTile);
    const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
    int64_t grid = (int64_t)ctx->num_sm * per_sm;       // persistent CTAs: T0 is staged once per CTA, not once per tile
    if (grid > ntile) grid = ntile;
    if (D == 2) {
        MP_CUDA(cudaFuncSetAttribute(k_reftensor<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_reftensor<2><<<(unsigned)grid, kRefThreads, smem, ctx->stream>>>(*mesh, kind, dir, L, lc, NK, d_T0, d_coef, coef_kind, scale, d_row_pos, d_elem_mat);
    } else {
        MP_CUDA(cudaFuncSetAttribute(k_reftensor<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_reftensor<3><<<(unsigned)grid, kRefThreads, smem, ctx->stream>>>(*mesh, kind, dir, L, lc, NK, d_T0, d_coef, coef_kind, scale, d_row_pos, d_elem_mat);
    }
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mp_elem_reftensor(mp_ctx *ctx, const mp_mesh *mesh, int kind, int dir, int lr, int lc, const double *d_T0,
                                 const double *d_coef, int coef_kind, double scale, double *d_elem_mat) {
    return reftensor_impl(ctx, mesh, kind, dir, lr, lc, d_T0, d_coef, coef_kind, scale, nullptr, d_elem_mat);
}

extern "C" int mp_elem_reftensor_rows(mp_ctx *ctx, const mp_mesh *mesh, int kind, int dir, int lr, int lc, const double *d_T0,
                                      const double *d_coef, int coef_kind, double scale, const int32_t *d_row_pos, double *d_elem_mat) {
    MP_CHECK(d_row_pos, "mp_elem_reftensor_rows: null row positions");
    return reftensor_impl(ctx, mesh, kind, dir, lr, lc, d_T0, d_coef, coef_kind, scale, d_row_pos, d_elem_mat);
}

extern "C" int mp_elem_p1_transport_quad(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const mp_quad *quad,
                                         const int32_t *d_vel_dofs, const double *d_vel0, const double *d_vel1, const double *d_vel2,
                                         const double *d_c0n, const double *d_c1n, const double *d_c2n, const double *d_c1new,
                                         const double *d_c2new, const double *d_fsrc, double *d_elem_mat, double *d_elem_vec) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(prm && quad && quad->nq > 0 && quad->nbv > 0 && quad->d_w && quad->d_lam && quad->d_vphi, "mp_elem_p1_transport_quad: bad quadrature");
    MP_CHECK(d_vel_dofs && d_vel0 && d_vel1 && (mesh->dim == 2 || d_vel2), "mp_elem_p1_transport_quad: velocity pointers");
    MP_CHECK(!d_elem_vec || d_c0n, "mp_elem_p1_transport_quad: rhs needs c0n");
    MP_CHECK(prm->supg != 5 || d_c1n, "mp_elem_p1_transport_quad: supg==5 needs c1n");
    MP_CHECK(!d_elem_vec || prm->ncomp < 2 || (d_c1n && d_c1new), "mp_elem_p1_transport_quad: ncomp>=2 needs c1n,c1new");
    MP_CHECK(!d_elem_vec || prm->ncomp < 3 || (d_c2n && d_c2new), "mp_elem_p1_transport_quad: ncomp==3 needs c2n,c2new");
    QuadView Q{quad->nq, quad->nbv, quad->d_w, quad->d_lam, quad->d_vphi};
    DISPATCH_DIM(mesh, k_p1_transport_quad, *prm, Q, d_vel_dofs, d_vel0, d_vel1, d_vel2 ? d_vel2 : d_vel1, d_c0n, d_c1n, d_c2n, d_c1new,
                 d_c2new, d_fsrc, d_elem_mat, d_elem_vec);
    return 0;
}

static int check_quad(const mp_quad *q, bool need_v, bool need_dv) {
    MP_CHECK(q && q->nq > 0 && q->d_w && q->d_lam, "quadrature tables missing");
    MP_CHECK(!need_v || (q->nbv > 0 && q->d_vphi), "quadrature: velocity basis table missing");
    MP_CHECK(!need_dv || q->d_vdphi, "quadrature: velocity basis gradient table missing");
    return 0;
}

extern "C" int mp_cell_fn(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, int fn, double dL, const double *d_phi, double *d_out) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_TRY(check_quad(quad, false, false));
    MP_CHECK(fn >= 0 && fn <= 2 && d_phi && d_out, "mp_cell_fn: bad argument");
    QuadView Q{quad->nq, quad->nbv, quad->d_w, quad->d_lam, quad->d_vphi};
    DISPATCH_DIM(mesh, k_cell_fn, Q, fn, dL, d_phi, d_out);
    return 0;
}

extern "C" int mp_elem_p1_mass_fn(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, int fn, double dL, double scale,
                                  const double *d_phi, double *d_elem_mat) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_TRY(check_quad(quad, false, false));
    MP_CHECK(fn >= 0 && fn <= 2 && d_phi && d_elem_mat, "mp_elem_p1_mass_fn: bad argument");
    QuadView Q{quad->nq, quad->nbv, quad->d_w, quad->d_lam, quad->d_vphi};
    DISPATCH_DIM(mesh, k_p1_mass_fn, Q, fn, dL, scale, d_phi, d_elem_mat);
    return 0;
}

extern "C" int mp_elem_compaction_rhs(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, const mp_compaction_params *prm,
                                      const double *d_phi, const double *d_gam, const double *d_buoy, double *d_evec_u0,
                                      double *d_evec_u1, double *d_evec_u2, double *d_evec_p) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_TRY(check_quad(quad, true, false));
    MP_CHECK(prm && d_phi && d_gam && d_buoy && d_evec_u0 && d_evec_u1 && (mesh->dim == 2 || d_evec_u2) && d_evec_p, "mp_elem_compaction_rhs: null argument");
    MP_CHECK(quad->nbv == (mesh->dim == 2 ? 6 : 10), "mp_elem_compaction_rhs: velocity basis must be P2");
    CompactionPrm P{prm->da, prm->R, prm->B, 2.0 * cos(prm->theta) - 1.0, prm->dL, {prm->zhat[0], prm->zhat[1], prm->zhat[2]}};
    QuadView Q{quad->nq, quad->nbv, quad->d_w, quad->d_lam, quad->d_vphi};
    DISPATCH_DIM(mesh, k_compaction_rhs, Q, P, d_phi, d_gam, d_buoy, d_evec_u0, d_evec_u1, d_evec_u2 ? d_evec_u2 : d_evec_u1, d_evec_p);
    return 0;
}

extern "C" int mp_elem_p1_porosity_quad(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, double da, double dt,
                                        const int32_t *d_vel_dofs, const double *d_vel0, const double *d_vel1, const double *d_vel2,
                                        const double *d_phi0, const double *d_gam, double *d_elem_mat, double *d_elem_vec) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_TRY(check_quad(quad, true, true));
    MP_CHECK(quad->nbv == (mesh->dim == 2 ? 6 : 10), "mp_elem_p1_porosity_quad: velocity basis must be P2");
    MP_CHECK(d_vel_dofs && d_vel0 && d_vel1 && (mesh->dim == 2 || d_vel2) && d_phi0 && d_gam, "mp_elem_p1_porosity_quad: null argument");
    QuadView Q{quad->nq, quad->nbv, quad->d_w, quad->d_lam, quad->d_vphi};
    DISPATCH_DIM(mesh, k_p1_porosity_quad, Q, quad->d_vdphi, da, dt, d_vel_dofs, d_vel0, d_vel1, d_vel2 ? d_vel2 : d_vel1, d_phi0, d_gam,
                 d_elem_mat, d_elem_vec);
    return 0;
}

extern "C" int mp_elem_p2_porosity_quad(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, double da, double dt, const int32_t *d_p2_dofs,
                                        const double *d_vel0, const double *d_vel1, const double *d_vel2, const double *d_phi0,
                                        const double *d_gam, double *d_elem_mat, double *d_elem_vec) {
    MP_TRY(check_mesh(ctx, mesh));
    MP_CHECK(quad && quad->d_w && quad->d_vphi && quad->d_vdphi && quad->nq > 0, "mp_elem_p2_porosity_quad: quadrature tables (incl. gradients) required");
    MP_CHECK(quad->nbv == (mesh->dim == 2 ? 6 : 10), "mp_elem_p2_porosity_quad: the tables must be those of the P2 basis");
    MP_CHECK(d_p2_dofs && d_vel0 && d_vel1 && (mesh->dim == 2 || d_vel2) && d_phi0 && d_gam, "mp_elem_p2_porosity_quad: null argument");
    MP_CHECK(d_elem_mat || d_elem_vec, "mp_elem_p2_porosity_quad: nothing to compute");
    QuadView Q{quad->nq, quad->nbv, quad->d_w, quad->d_lam, quad->d_vphi};
    const unsigned grid = (unsigned)cdiv(mesh->ncell * 32, 256);
    if (grid == 0) return 0;
    if (mesh->dim == 2) k_p2_porosity_quad<2><<<grid, 256, 0, ctx->stream>>>(*mesh, Q, quad->d_vdphi, da, dt, d_p2_dofs, d_vel0, d_vel1, d_vel1, d_phi0, d_gam, d_elem_mat, d_elem_vec);
    else k_p2_porosity_quad<3><<<grid, 256, 0, ctx->stream>>>(*mesh, Q, quad->d_vdphi, da, dt, d_p2_dofs, d_vel0, d_vel1, d_vel2, d_phi0, d_gam, d_elem_mat, d_elem_vec);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    return 0;
}
