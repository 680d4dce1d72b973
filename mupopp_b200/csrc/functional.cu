// Scalar functionals on the device: boundary fluxes, volume integrals and L2 error norms.
//
// Replaces dolfin.assemble(<scalar form>) and errornorm() that the reference scripts call every output step:
//   flux1 = assemble(c00*phi0*dot(u0, n)*ds(1))            /root/reference/src/run/CCS2D/CCS_3comp.py:277
//   assemble(c00*dot(u, unitNormal)*ds(1)), assemble(c02*dx), assemble(CaCO3_frac*Da0*c00*c01*dx)
//                                                           /root/reference/src/run/stokes_ADR_3D/master_advective_stokes.py:232-265
//   errornorm(u, uh, 'L2')                                  /root/reference/src/benchmark/sphere_3D_pure_shear/sphere.py:157-159
// One thread per entity (cell or boundary facet): gather the local dofs of every factor, evaluate the product at the points of
// an exact quadrature rule (basis tables prepared once on the host per reference entity), accumulate, and reduce with the
// deterministic last-block scheme of reduce.cuh (no floating-point atomics; allreduced over ranks like every other scalar).
// Doing this on the device removes the device->host copy of whole fields that a CPU-side diagnostic would force every step.
#include "reduce.cuh"

namespace {

constexpr int kMaxNb = 20;   // P3 on tetrahedra

__global__ void __launch_bounds__(kThreads)
k_integrate(int dim, mp_functional F, int64_t nent, const int32_t *__restrict__ ent_cell, const int32_t *__restrict__ ent_tab,
            const double *__restrict__ ent_weight, const double *__restrict__ ent_g, double g0, double g1, double g2,
            double *partials, mp::ReduceCtl *ctl, double *scal) {
    double acc[1] = {0.0};
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nent; e += gridDim.x * (int64_t)blockDim.x) {
        const int64_t c = ent_cell[e];
        const int t = ent_tab ? ent_tab[e] : 0;
        double g[3] = {g0, g1, g2};
        if (ent_g)
            for (int a = 0; a < dim; ++a) g[a] = ent_g[e * dim + a];
        // the vector factor, projected on g, as ONE set of local coefficients (Lagrange) or one number (cellwise constant)
        double vloc[kMaxNb];
        double vconst = 1.0;
        if (F.vec_kind == 1) {
            for (int b = 0; b < F.nbv; ++b) {
                const int32_t d = F.d_vdofs[c * F.nbv + b];
                double s = 0.0;
                for (int a = 0; a < dim; ++a) s += F.d_vec[a][d] * g[a];
                vloc[b] = s;
            }
        } else if (F.vec_kind == 2) {
            vconst = 0.0;
            for (int a = 0; a < dim; ++a) vconst += F.d_vec[0][c * dim + a] * g[a];
        }
        double sum = 0.0;
        for (int q = 0; q < F.nq; ++q) {
            double v = F.d_w[q] * vconst;
            if (F.vec_kind == 1) {
                const double *tab = F.d_vtab + ((int64_t)t * F.nq + q) * F.nbv;
                double s = 0.0;
                for (int b = 0; b < F.nbv; ++b) s += tab[b] * vloc[b];
                v *= s;
            }
            for (int k = 0; k < F.nfield; ++k) {
                const int nb = F.nb[k];
                const double *tab = F.d_tab[k] + ((int64_t)t * F.nq + q) * nb;
                const int32_t *dofs = F.d_dofs[k] + c * nb;
                const double *f = F.d_field[k];
                double s = 0.0;
                for (int b = 0; b < nb; ++b) s += tab[b] * f[dofs[b]];
                v *= s;
            }
            sum += v;
        }
        acc[0] += ent_weight[e] * sum;
    }
    const int sl_[1] = {S_TMP0};
    grid_reduce<1>(acc, partials, ctl, scal, sl_, false);
}

// max over cells of the cell average of |u|: core.u_max (/root/reference/src/modules/core.py:23-39) behind compute_dt (:43-48).
// vec_kind 1: Lagrange components tabulated at the quadrature points; 2: one constant vector per cell (mode F).  The maximum of
// non-negative doubles is taken on their bit patterns (order independent, hence deterministic).
__global__ void __launch_bounds__(kThreads)
k_max_cell_speed(int dim, int64_t ncell, int vec_kind, const double *__restrict__ u0, const double *__restrict__ u1, const double *__restrict__ u2,
                 const int32_t *__restrict__ vdofs, int nbv, const double *__restrict__ vtab, int nq, const double *__restrict__ w,
                 unsigned long long *out) {
    double best = 0.0;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * (int64_t)blockDim.x) {
        double mean = 0.0;
        if (vec_kind == 2) {
            double s = 0.0;
            for (int a = 0; a < dim; ++a) s += u0[c * dim + a] * u0[c * dim + a];
            mean = sqrt(s);
        } else {
            for (int q = 0; q < nq; ++q) {
                double v[3] = {0.0, 0.0, 0.0};
                for (int b = 0; b < nbv; ++b) {
                    const int32_t d = vdofs[c * nbv + b];
                    const double t = vtab[q * nbv + b];
                    v[0] += t * u0[d];
                    v[1] += t * u1[d];
                    if (dim == 3) v[2] += t * u2[d];
                }
                mean += w[q] * sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            }
        }
        best = fmax(best, mean);
    }
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(best));
}

}  // namespace

extern "C" int mp_max_cell_speed(mp_ctx *ctx, int32_t dim, int64_t ncell, int32_t vec_kind, const double *d_u0, const double *d_u1,
                                 const double *d_u2, const int32_t *d_vdofs, int32_t nbv, const double *d_vtab, int32_t nq, const double *d_w,
                                 double *result) {
    MP_CHECK(ctx && result && d_u0 && (dim == 2 || dim == 3) && ncell >= 0, "mp_max_cell_speed: bad argument");
    MP_CHECK(vec_kind == 2 || (vec_kind == 1 && d_u1 && (dim == 2 || d_u2) && d_vdofs && d_vtab && d_w && nbv > 0 && nq > 0),
             "mp_max_cell_speed: bad vector description");
    unsigned long long *d_out = reinterpret_cast<unsigned long long *>(ctx->d_scalars + 62);     // scratch at the end of the scalar block
    MP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), ctx->stream));
    if (ncell > 0) {
        int64_t g = cdiv(ncell, kThreads);
        const int64_t cap = (int64_t)ctx->num_sm * 8;
        if (g > cap) g = cap;
        k_max_cell_speed<<<(unsigned)g, kThreads, 0, ctx->stream>>>(dim, ncell, vec_kind, d_u0, d_u1, d_u2, d_vdofs, nbv, d_vtab, nq, d_w, d_out);
        mp::count_launch(ctx);
        MP_CUDA(cudaGetLastError());
    }
    double h = 0.0;
    MP_CUDA(cudaMemcpyAsync(&h, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    MP_TRY(mp::allreduce_max_host(ctx, &h));
    *result = h;
    return 0;
}

extern "C" int mp_integrate(mp_ctx *ctx, int32_t dim, const mp_functional *f, int64_t nent, const int32_t *d_ent_cell,
                            const int32_t *d_ent_tab, const double *d_ent_weight, const double *d_ent_g, const double *g_const,
                            double *result) {
    MP_CHECK(ctx && f && result && nent >= 0, "mp_integrate: bad argument");
    MP_CHECK(dim == 2 || dim == 3, "mp_integrate: dim must be 2 or 3");
    MP_CHECK(f->nfield >= 0 && f->nfield <= 4 && f->nq > 0 && f->d_w, "mp_integrate: bad functional");
    MP_CHECK(f->vec_kind >= 0 && f->vec_kind <= 2 && (f->vec_kind != 1 || (f->nbv > 0 && f->nbv <= kMaxNb && f->d_vdofs && f->d_vtab)),
             "mp_integrate: bad vector factor");
    MP_CHECK(f->vec_kind == 0 || d_ent_g || g_const, "mp_integrate: a vector factor needs per-entity normals or a constant direction");
    for (int k = 0; k < f->nfield; ++k)
        MP_CHECK(f->d_field[k] && f->d_dofs[k] && f->d_tab[k] && f->nb[k] > 0 && f->nb[k] <= kMaxNb, "mp_integrate: bad scalar factor %d", k);
    MP_CHECK(nent == 0 || (d_ent_cell && d_ent_weight), "mp_integrate: null entity arrays");
    int64_t g = cdiv(nent > 0 ? nent : 1, kThreads);
    int64_t cap = (int64_t)ctx->num_sm * 8;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (g > cap) g = cap;
    const double g0 = g_const ? g_const[0] : 0.0, g1 = g_const ? g_const[1] : 0.0, g2 = (g_const && dim == 3) ? g_const[2] : 0.0;
    k_integrate<<<(unsigned)g, kThreads, 0, ctx->stream>>>(dim, *f, nent, d_ent_cell, d_ent_tab, d_ent_weight, d_ent_g, g0, g1, g2,
                                                           ctx->d_partials, ctx->d_ctl, ctx->d_scalars);
    mp::count_launch(ctx);
    MP_CUDA(cudaGetLastError());
    MP_TRY(mp::reduce_slots_public(ctx, S_TMP0, 1));
    MP_CUDA(cudaMemcpyAsync(ctx->h_scalars + S_TMP0, ctx->d_scalars + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MP_CUDA(cudaStreamSynchronize(ctx->stream));
    *result = ctx->h_scalars[S_TMP0];
    return 0;
}
