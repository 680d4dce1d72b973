"""Thin object layer over the C-ABI: context, device mesh, scatter plans, CSR operators, solvers.

torch owns every device buffer; the numerics are the hand-written sm_100a kernels of
libmupopp_b200.so.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import MpCsr, MpMesh, MpSolveInfo, MpTransportParams, check, ptr


class Context:
    """One CUDA device + stream (+ optional NCCL communicator)."""

    def __init__(self, device=0, use_torch_stream=True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MpError("mupopp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream if use_torch_stream else None
        h = C.c_void_p()
        check(self.lib.mp_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)), "mp_ctx_create")
        self.h = h
        self.nranks, self.rank = 1, 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.mp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.mp_ctx_sync(self.h), "mp_ctx_sync")

    def launch_count(self):
        return int(self.lib.mp_ctx_launch_count(self.h))

    def set_option(self, name, value):
        check(self.lib.mp_ctx_set_option(self.h, name.encode(), int(value)), "mp_ctx_set_option")

    def timing(self, enable, max_samples=4096):
        check(self.lib.mp_ctx_timing(self.h, 1 if enable else 0, max_samples), "mp_ctx_timing")

    def timing_read(self):
        n, ms = C.c_int(), C.c_double()
        check(self.lib.mp_ctx_timing_read(self.h, C.byref(n), C.byref(ms)), "mp_ctx_timing_read")
        return n.value, ms.value

    # ---- tensors
    def zeros(self, n, dtype=torch.float64):
        return torch.zeros(int(n), dtype=dtype, device=self.device)

    def empty(self, n, dtype=torch.float64):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def to_device(self, a, dtype=None):
        t = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device)

    # ---- multi-GPU
    def init_comm(self, nranks, rank, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self.lib.mp_comm_init(self.h, nranks, rank, buf), "mp_comm_init")
        self.nranks, self.rank = nranks, rank

    def enable_peer(self, allgather=None):
        """Attach the peer-memory (CUDA IPC over NVLink) mailbox windows: the Krylov inner products are then allreduced inside
        the reducing kernels.  `allgather(obj) -> [obj of rank 0, ...]` is the host-side exchange of the 64-byte IPC handles
        (default: torch.distributed.all_gather_object).  Returns True when the fast path is on."""
        if allgather is None:
            allgather = _dist_allgather
        if self.nranks == 1:
            allgather = lambda obj: [obj]
        buf = (C.c_uint8 * 64)()
        check(self.lib.mp_comm_peer_export(self.h, buf), "mp_comm_peer_export")
        handles = allgather(bytes(buf))
        if len(handles) != self.nranks:
            raise _lib.MpError("enable_peer: %d handles for %d ranks" % (len(handles), self.nranks))
        blob = (C.c_uint8 * (64 * self.nranks)).from_buffer_copy(b"".join(handles))
        err = None
        try:
            check(self.lib.mp_comm_peer_attach(self.h, blob), "mp_comm_peer_attach")
        except _lib.MpError as e:          # e.g. CUDA IPC not permitted between these processes: every rank must then stay on NCCL
            err = str(e)
        errs = allgather(err)
        self._allgather = allgather
        if any(e is not None for e in errs):
            if err is None:
                self.set_option("peer", 0)
            self.peer_error = next(e for e in errs if e is not None)
            return False
        return self.peer_enabled()

    def peer_enabled(self):
        return bool(self.lib.mp_comm_peer_enabled(self.h))

    def unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        check(self.lib.mp_nccl_unique_id(buf), "mp_nccl_unique_id")
        return bytes(buf)

    # ---- BLAS-1
    def dot(self, x, y):
        r = C.c_double()
        check(self.lib.mp_dot(self.h, x.numel(), ptr(x), ptr(y), C.byref(r)), "mp_dot")
        return r.value

    def vmul(self, a, x, y, z):
        check(self.lib.mp_vmul(self.h, z.numel(), a, ptr(x), ptr(y), ptr(z)), "mp_vmul")

    def axpby(self, a, x, b, y):
        check(self.lib.mp_axpby(self.h, y.numel(), a, ptr(x), b, ptr(y)), "mp_axpby")


class DeviceMesh:
    """Mesh on the device.  `vert2dof` (int array, one P1 dof per vertex) identifies vertices of a periodic space; then
    `cell_dofs` differs from `cells`, `ndof < nvert`, and nodal fields live in dof numbering."""

    def __init__(self, ctx: Context, mesh, vert2dof=None):
        self.ctx = ctx
        self.host = mesh
        self.dim = mesh.dim
        self.ncell = mesh.num_cells()
        self.nvert = mesh.num_vertices()
        self.cells = ctx.to_device(mesh.cells, torch.int32)
        self.coords = ctx.to_device(mesh.coords, torch.float64)
        if vert2dof is None:
            self.cell_dofs, self.ndof, dp = self.cells, self.nvert, None
        else:
            v2d = np.asarray(vert2dof, dtype=np.int32)
            self.cell_dofs = ctx.to_device(np.ascontiguousarray(v2d[mesh.cells]), torch.int32)
            self.ndof = int(v2d.max()) + 1
            dp = self.cell_dofs.data_ptr()
        self.c = MpMesh(self.dim, self.ncell, self.nvert, self.cells.data_ptr(), self.coords.data_ptr(), dp, None)

    def use_row_ordered_storage(self, plan):
        """Make the P1 element kernels write their output in the row order of `plan` (a P1xP1 plan of this mesh's dofmap), so
        that the scatter kernels stream it.  Every element tensor produced with this mesh must then be scattered through `plan`."""
        pos = C.c_void_p()
        check(self.ctx.lib.mp_plan_row_positions(plan.h, C.byref(pos)), "mp_plan_row_positions")
        self.c.d_row_pos = pos.value
        plan.ordered = True


class Plan:
    """Scatter plan for one (row dofmap, column dofmap) pair; see include/mupopp_b200.h."""

    def __init__(self, ctx: Context, row_dofs, col_dofs, nrow, ncol):
        self.ctx = ctx
        self.row_dofs = row_dofs            # keep alive
        self.col_dofs = col_dofs
        ncell, self.lr = row_dofs.shape
        self.lc = col_dofs.shape[1] if col_dofs is not None else 0
        self.ncell, self.nrow, self.ncol = ncell, int(nrow), int(ncol)
        h = C.c_void_p()
        check(ctx.lib.mp_plan_create(ctx.h, ptr(row_dofs), self.lr, ptr(col_dofs), self.lc, ncell, nrow, ncol, C.byref(h)),
              "mp_plan_create")
        self.h = h
        self.ordered = False
        self.nnz = int(ctx.lib.mp_plan_nnz(h)) if self.lc else 0
        self.max_row_nnz = int(ctx.lib.mp_plan_max_row_nnz(h)) if self.lc else 0
        self._csr = None

    def __del__(self):
        try:
            if self.h:
                self.ctx.lib.mp_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def csr_pattern(self):
        if self._csr is None:
            rowptr = self.ctx.empty(self.nrow + 1, torch.int32)
            colidx = self.ctx.empty(max(self.nnz, 1), torch.int32)
            check(self.ctx.lib.mp_plan_export_csr(self.ctx.h, self.h, ptr(rowptr), ptr(colidx)), "mp_plan_export_csr")
            self._csr = (rowptr, colidx[: self.nnz])
        return self._csr

    def sellu_info(self):
        """(number of (slice, column-position) pairs that need explicit columns, or -1; meta pointer; explicit-column pointer)."""
        nx, me, xc = C.c_int64(), C.c_void_p(), C.c_void_p()
        check(self.ctx.lib.mp_plan_sellu_info(self.h, C.byref(nx), C.byref(me), C.byref(xc)), "mp_plan_sellu_info")
        return nx.value, me.value, xc.value

    def sell_info(self):
        ns, ne, sp, sc = C.c_int64(), C.c_int64(), C.c_void_p(), C.c_void_p()
        check(self.ctx.lib.mp_plan_sell_info(self.h, C.byref(ns), C.byref(ne), C.byref(sp), C.byref(sc)), "mp_plan_sell_info")
        return ns.value, ne.value, sp.value, sc.value

    def new_matrix(self, sell=True):
        rowptr, colidx = self.csr_pattern()
        return CsrMatrix(self.ctx, self.nrow, self.ncol, rowptr, colidx, self.ctx.zeros(self.nnz), self.max_row_nnz,
                         plan=self if sell else None)

    def row_positions(self):
        """Device pointer of pos[ncell*lr]: incidence index of (cell, local row) -- where a row-ordered element kernel stores that row."""
        pos = C.c_void_p()
        check(self.ctx.lib.mp_plan_row_positions(self.h, C.byref(pos)), "mp_plan_row_positions")
        return pos

    def reftensor_scatter(self, dm, kind, direction, lr, lc, T0, coef, coef_kind, scale, emat, A):
        """One reference-tensor block straight into the CSR matrix A: element kernel with row-ordered output (mp_elem_reftensor_rows)
        followed by the streaming scatter -- the P2/P3 blocks of mode R, compaction and Stokes."""
        check(self.ctx.lib.mp_elem_reftensor_rows(self.ctx.h, C.byref(dm.c), int(kind), int(direction), int(lr), int(lc), ptr(T0), ptr(coef),
                                                  int(coef_kind), float(scale), self.row_positions(), ptr(emat)), "mp_elem_reftensor_rows")
        self.scatter_matrix(emat, A, ordered=True)

    def scatter_matrix(self, elem_mat, A, ordered=None):
        ordered = self.ordered if ordered is None else ordered
        fn = self.ctx.lib.mp_scatter_matrix_ordered if ordered else self.ctx.lib.mp_scatter_matrix
        check(fn(self.ctx.h, self.h, ptr(elem_mat), ptr(A.vals)), "mp_scatter_matrix")
        A.sell_dirty = True

    def scatter_vector(self, elem_vec, out, beta=0.0):
        fn = self.ctx.lib.mp_scatter_vector_ordered if self.ordered else self.ctx.lib.mp_scatter_vector
        check(fn(self.ctx.h, self.h, ptr(elem_vec), beta, ptr(out)), "mp_scatter_vector")


class CsrMatrix:
    def __init__(self, ctx, nrow, ncol, rowptr, colidx, vals, max_row_nnz=0, plan=None):
        self.ctx = ctx
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.nnz = int(vals.numel())
        self.plan = plan
        self.sell_vals, self.sell_dirty = None, True
        sp = sc = sv = me = xc = None
        self.sell_entries, self.sell_explicit = 0, -1
        if plan is not None and plan.lc > 0:
            _, nent, sp, sc = plan.sell_info()
            self.sell_vals = ctx.zeros(max(nent, 1))
            sv = self.sell_vals.data_ptr()
            self.sell_entries = int(nent)
            self.sell_explicit, me, xc = plan.sellu_info()
        self.c = MpCsr(self.nrow, self.ncol, self.nnz, int(max_row_nnz), 0, rowptr.data_ptr(), colidx.data_ptr(), vals.data_ptr(),
                       sp, sc, sv, me, xc)
        self.dinv = None

    def sync_sell(self):
        """Refresh the SELL-32 mirror used by SpMV after the CSR values changed (assembly / Dirichlet)."""
        if self.sell_vals is not None and self.sell_dirty:
            check(self.ctx.lib.mp_csr_to_sell(self.ctx.h, self.plan.h, ptr(self.vals), ptr(self.sell_vals)), "mp_csr_to_sell")
            self.sell_dirty = False

    def spmv(self, x, y):
        self.sync_sell()
        check(self.ctx.lib.mp_spmv(self.ctx.h, C.byref(self.c), ptr(x), ptr(y)), "mp_spmv")
        return y

    def jacobi_setup(self):
        if self.dinv is None:
            self.dinv = self.ctx.empty(self.nrow)
        check(self.ctx.lib.mp_jacobi_setup(self.ctx.h, C.byref(self.c), ptr(self.dinv)), "mp_jacobi_setup")
        return self.dinv

    def line_setup(self, stride, spd=True):
        """Factor the line (block-tridiagonal) Jacobi preconditioner: rows i, i + stride, ... form one line (mp_line_setup).
        `spd`: insist on positive pivots (PCG); False for the non-symmetric transport row.  The three factor arrays are reused."""
        stride = int(stride)
        old = getattr(self, "_line", None)
        lo, invw, mb = old[1:] if old is not None else (self.ctx.empty(self.nrow) for _ in range(3))
        check(self.ctx.lib.mp_line_setup(self.ctx.h, C.byref(self.c), stride, int(bool(spd)), ptr(lo), ptr(invw), ptr(mb)), "mp_line_setup")
        self._line = (stride, lo, invw, mb)
        return self._line

    def line_apply(self, r, z):
        st, lo, invw, mb = self._line
        work = self.ctx.empty(self.nrow)
        check(self.ctx.lib.mp_line_apply(self.ctx.h, self.nrow, st, ptr(lo), ptr(invw), ptr(mb), ptr(r), ptr(z), ptr(work)), "mp_line_apply")
        return z

    def apply_dirichlet(self, rhs, isbc, g, matrix=True):
        check(self.ctx.lib.mp_apply_dirichlet_sym(self.ctx.h, self.nrow, ptr(self.rowptr), ptr(self.colidx),
                                                  ptr(self.vals) if matrix else None, ptr(rhs), ptr(isbc), ptr(g)),
              "mp_apply_dirichlet_sym")
        if matrix:
            self.sell_dirty = True

    def to_scipy(self):
        """Test helper: copy to host as scipy CSR."""
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals.cpu().numpy(), self.colidx.cpu().numpy(), self.rowptr.cpu().numpy()),
                             shape=(self.nrow, self.ncol))

    def solve(self, method, b, x, rtol=1e-12, atol=0.0, maxit=10000, halo=None, check_every=0, refresh_jacobi=True, restart=50,
              cheb_degree=3, cheb_ratio=30.0, line_stride=0):
        if method in ("pcg_line", "bicgstab_line"):
            if getattr(self, "_line", None) is None or self._line[0] != int(line_stride) or refresh_jacobi:
                self.line_setup(line_stride, spd=method == "pcg_line")
            self.sync_sell()
            info = MpSolveInfo(rtol, atol, int(maxit), int(check_every), 0, 0, 0.0)
            st, lo, invw, mb = self._line
            fn = self.ctx.lib.mp_solve_pcg_line if method == "pcg_line" else self.ctx.lib.mp_solve_bicgstab_line
            check(fn(self.ctx.h, C.byref(self.c), st, ptr(lo), ptr(invw), ptr(mb), ptr(b), ptr(x),
                     halo.h if halo is not None else None, C.byref(info)), "mp_solve_" + method)
            return SolveResult(info.niter, bool(info.converged), info.relres)
        if self.dinv is None or refresh_jacobi:
            self.jacobi_setup()
        self.sync_sell()
        info = MpSolveInfo(rtol, atol, int(maxit), int(check_every), 0, 0, 0.0)
        hh = halo.h if halo is not None else None
        if method == "pcg_chebyshev":
            if getattr(self, "_lmax", None) is None or refresh_jacobi:
                lm = C.c_double()
                check(self.ctx.lib.mp_spectral_bound(self.ctx.h, C.byref(self.c), ptr(self.dinv), C.byref(lm)), "mp_spectral_bound")
                self._lmax = lm.value
            check(self.ctx.lib.mp_solve_pcg_chebyshev(self.ctx.h, C.byref(self.c), ptr(self.dinv), ptr(b), ptr(x), hh, int(cheb_degree),
                                                      self._lmax / float(cheb_ratio), self._lmax, C.byref(info)), "mp_solve_pcg_chebyshev")
        elif method == "gmres":
            check(self.ctx.lib.mp_solve_gmres(self.ctx.h, C.byref(self.c), ptr(self.dinv), ptr(b), ptr(x), hh, int(restart),
                                              C.byref(info)), "mp_solve_gmres")
        else:
            fn = {"pcg": self.ctx.lib.mp_solve_pcg, "cg": self.ctx.lib.mp_solve_pcg,
                  "bicgstab": self.ctx.lib.mp_solve_bicgstab}[method]
            check(fn(self.ctx.h, C.byref(self.c), ptr(self.dinv), ptr(b), ptr(x), hh, C.byref(info)), "mp_solve_" + method)
        return SolveResult(info.niter, bool(info.converged), info.relres)


class SolveResult:
    def __init__(self, niter, converged, relres):
        self.niter, self.converged, self.relres = niter, converged, relres

    def __repr__(self):
        return "SolveResult(niter=%d, converged=%s, relres=%.3e)" % (self.niter, self.converged, self.relres)


class Halo:
    """Ghost-dof exchange plan (NCCL send/recv).  Ghost entries are contiguous per neighbour."""

    def __init__(self, ctx: Context, neigh_rank, send_lists, recv_start, recv_count):
        self.ctx = ctx
        n = len(neigh_rank)
        send_off = np.zeros(n + 1, dtype=np.int32)
        for k, s in enumerate(send_lists):
            send_off[k + 1] = send_off[k] + len(s)
        send_idx = np.concatenate(send_lists).astype(np.int32) if n and send_off[-1] > 0 else np.zeros(1, dtype=np.int32)
        self.d_send_idx = ctx.to_device(send_idx, torch.int32)
        arr = lambda a: (C.c_int32 * max(len(a), 1))(*[int(v) for v in a])
        h = C.c_void_p()
        check(ctx.lib.mp_halo_create(ctx.h, n, arr(neigh_rank), arr(send_off), ptr(self.d_send_idx), arr(recv_start),
                                     arr(recv_count), C.byref(h)), "mp_halo_create")
        self.h = h
        self.neigh_rank = [int(r) for r in neigh_rank]
        self.recv_start = [int(r) for r in recv_start]
        self.recv_count = [int(r) for r in recv_count]

    def exchange(self, x):
        check(self.ctx.lib.mp_halo_exchange(self.ctx.h, self.h, ptr(x)), "mp_halo_exchange")

    def enable_peer(self, n_owned, n_local, allgather=None):
        """Ghost window of this halo in peer memory: the SELL SpMV of the Krylov loops then pushes the interface values into the
        neighbours' windows itself (no k_pack + ncclSend/ncclRecv).  Collective over all ranks."""
        ctx = self.ctx
        if allgather is None:
            allgather = getattr(ctx, "_allgather", None) or _dist_allgather
        nghost = int(n_local) - int(n_owned)
        buf = (C.c_uint8 * 64)()
        check(ctx.lib.mp_halo_peer_export(ctx.h, self.h, int(n_owned), nghost, buf), "mp_halo_peer_export")
        mine = dict(handle=bytes(buf), neigh=self.neigh_rank, goff=[s - int(n_owned) for s in self.recv_start], nghost=nghost)
        everyone = allgather(mine)
        n = len(self.neigh_rank)
        handles, goff, gcap, slot = b"", [], [], []
        for q in self.neigh_rank:
            theirs = everyone[q]
            j = theirs["neigh"].index(ctx.rank)        # the relation is symmetric for owner-computes partitions
            handles += theirs["handle"]
            goff.append(theirs["goff"][j])
            gcap.append(theirs["nghost"])
            slot.append(j)
        hb = (C.c_uint8 * max(64 * n, 1)).from_buffer_copy(handles if n else b"\0")
        a64 = lambda a: (C.c_int64 * max(len(a), 1))(*[int(v) for v in a])
        a32 = lambda a: (C.c_int32 * max(len(a), 1))(*[int(v) for v in a])
        check(ctx.lib.mp_halo_peer_attach(ctx.h, self.h, hb, a64(goff), a64(gcap), a32(slot)), "mp_halo_peer_attach")
        return True

    def __del__(self):
        try:
            if self.h:
                self.ctx.lib.mp_halo_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _dist_allgather(obj):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def transport_params(Pe, Da, phi, alpha, dt, sigma, rho0, gamma, fracs, zhat, K, reaction, ncomp, supg):
    p = MpTransportParams()
    p.Pe, p.Da, p.phi, p.beta, p.dt = float(Pe), float(Da), float(phi), float(alpha) / float(phi), float(dt)
    p.sigma, p.rho0, p.gamma = float(sigma), float(rho0), float(gamma)
    for i in range(3):
        p.frac[i] = float(fracs[i])
        p.zhat[i] = float(zhat[i]) if i < len(zhat) else 0.0
    p.Kconst = float(K) if np.isscalar(K) else 0.0
    p.reaction = {"c0c1": 0, "c0c1sq": 1}[reaction]
    p.ncomp, p.supg, p.pad = int(ncomp), int(supg), 0
    return p
