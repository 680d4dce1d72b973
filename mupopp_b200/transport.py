"""Mode-F timestep driver: the segregated DarcyAdvection / CCS step on P1 fields, all numerics on the GPU.

One step = the block-triangular sequence that is algebraically what the reference's monolithic
`solve(a==L, sol, bc)` does (/root/reference/src/run/CCS2D/CCS_3comp.py:257-263):
    1. c1, c2 consistent-mass solves (Jacobi-PCG)         mupopp.py:826, :936-937, :1234-1235
    2. c0 Crank-Nicolson + SUPG ADR solve (Jacobi-BiCGStab) with the lagged velocity   mupopp.py:823-848
    3. pressure solve div(k grad p) (Jacobi-PCG), warm-started from p^n
    4. fused cellwise velocity recovery u = -(K/phi)(grad p + sigma rho zhat)
Mode F is the primal P1 form named by BASELINE.json's north_star (DG0 velocity), see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from ._lib import check, ptr
from .device import Context, DeviceMesh, Plan, transport_params


@dataclass
class TransportModel:
    """Physical parameters of the monolithic DarcyAdvection / CCS forms (names follow mupopp.py)."""
    Pe: float = 100.0
    Da: float = 10.0
    phi: float = 0.01
    alpha: float = 0.005
    dt: float = 0.01
    K: object = 1.0                  # float or P1 nodal array
    sigma: float = -1.0              # -1: DarcyAdvection (mupopp.py:807), +1: CCS (mupopp.py:1204)
    rho0: float = 1.0
    gamma: float = 1.0
    reaction: str = "c0c1sq"
    fracs: tuple = (0.75, 1.0, 0.0)
    ncomp: int = 2
    SUPG: int = 1
    zhat: tuple = (0.0, 1.0)

    @staticmethod
    def darcy_two_component(Pe=100, Da=10.0, phi=0.01, alpha=0.005, dt=1e-2, K=1.0, zh=(0.0, 1.0), SUPG=1, gam_rho=-1.0, rho_0=1.0):
        fac = (24.0 + 12.0 + 3.0 * 16.0) / 2.0 / 56.0                        # mupopp.py:820
        return TransportModel(Pe, Da, phi, alpha, dt, K, -1.0, rho_0, -gam_rho, "c0c1sq", (fac, 1.0, 0.0), 2, SUPG, tuple(zh))

    @staticmethod
    def darcy_three_component(Pe=100, Da=10.0, phi=0.01, alpha=0.005, dt=1e-2, K=1.0, zh=(0.0, 1.0), SUPG=1, gam_rho=-1.0, rho_0=1.0):
        T = 392.32                                                           # mupopp.py:925-931
        return TransportModel(Pe, Da, phi, alpha, dt, K, -1.0, rho_0, -gam_rho, "c0c1sq", (84.0 / T, 112.0 / T, 12.0 / T), 3, SUPG, tuple(zh))

    @staticmethod
    def ccs_three_component(Pe=100, Da=10.0, phi=0.01, alpha=0.005, dt=1e-2, K=1.0, zh=(0.0, 1.0), SUPG=1, gam_rho=1.0):
        T = 698.0                                                            # mupopp.py:1217-1229
        return TransportModel(Pe, Da, phi, alpha, dt, K, 1.0, 1.0, gam_rho, "c0c1", (62.0 / T, 278.0 / T, 100.0 / T), 3, SUPG, tuple(zh))


@dataclass
class BoundaryDataF:
    p_isbc: np.ndarray            # (V,) uint8: p = 0 (natural boundary of the mixed form)
    flux_load: np.ndarray         # (V,) int_{Gamma_u} (u_D.n) q_i ds
    c0_isbc: np.ndarray
    c0_g: np.ndarray


def _eval_marker(marker, x):
    """marker: callable on an (N,d) array returning a bool array, or a pointwise callable x->bool."""
    try:
        r = np.asarray(marker(x))
        if r.shape == (x.shape[0],):
            return r.astype(bool)
    except Exception:
        pass
    return np.array([bool(marker(p)) for p in x])


def _eval_value(value, x, ncomp):
    if callable(value):
        try:
            r = np.asarray(value(x), dtype=float)
            if r.shape == (x.shape[0], ncomp) or (ncomp == 1 and r.shape == (x.shape[0],)):
                return r.reshape(x.shape[0], ncomp)
        except Exception:
            pass
        return np.array([np.atleast_1d(np.asarray(value(p), dtype=float)) for p in x]).reshape(x.shape[0], ncomp)
    return np.broadcast_to(np.atleast_1d(np.asarray(value, dtype=float)), (x.shape[0], ncomp)).copy()


def make_boundary_data(mesh, u_marker, u_value, c0_marker=None, c0_value=0.0, periodic=None):
    """Host setup of the Dirichlet/Neumann data (DOLFIN 'topological' facet marking: a facet is on Gamma_u if
    all its vertices and its midpoint satisfy the marker).  One-time; not part of the timestep.
    `periodic`: a mesh.PeriodicMap -- its faces are not boundary and the result is returned in dof numbering."""
    fv, n, meas, _ = mesh.boundary_facets(periodic.axes if periodic is not None else ())
    d = mesh.dim
    V = mesh.num_vertices()
    bverts = np.unique(fv.ravel())
    vin = np.zeros(V, dtype=bool)
    vin[bverts] = _eval_marker(u_marker, mesh.coords[bverts])
    inside = vin[fv].all(1)
    if inside.any():
        idx = np.nonzero(inside)[0]
        inside[idx] &= _eval_marker(u_marker, mesh.coords[fv[idx]].mean(1))
    flux = np.zeros(V)
    idx = np.nonzero(inside)[0]
    if idx.size:
        mloc = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
        xf = mesh.coords[fv[idx]].reshape(-1, d)
        uD = _eval_value(u_value, xf, d).reshape(idx.size, d, d)                  # (F, vertex, comp)
        g = np.einsum("fvc,fc->fv", uD, n[idx])
        contrib = meas[idx][:, None] * (g @ mloc.T)
        np.add.at(flux, fv[idx].ravel(), contrib.ravel())
    p_isbc = np.zeros(V, dtype=np.uint8)
    p_isbc[np.unique(fv[~inside].ravel())] = 1
    c0_isbc = np.zeros(V, dtype=np.uint8)
    c0_g = np.zeros(V)
    if c0_marker is not None:
        sel = bverts[_eval_marker(c0_marker, mesh.coords[bverts])]
        c0_isbc[sel] = 1
        c0_g[sel] = _eval_value(c0_value, mesh.coords[sel], 1)[:, 0]
    if periodic is not None:
        c0m = periodic.fold_any(c0_isbc)
        g = np.zeros(periodic.ndof)
        g[periodic.vert2dof[c0_isbc.astype(bool)]] = c0_g[c0_isbc.astype(bool)]
        return BoundaryDataF(periodic.fold_any(p_isbc), periodic.fold_sum(flux), c0m, g)
    return BoundaryDataF(p_isbc, flux, c0_isbc, c0_g)


def lattice_line_stride(mesh, vert2dof=None, n_owned=None):
    """Rows per lattice plane of a structured mesh (`mesh.grid`): the dofs of the grid line through in-plane position l are
    l, l + stride, l + 2 stride, ...  (BoxMesh numbers vertices iz*npl + iy*(nx+1) + ix, periodic x/y faces leave nx*ny dofs per
    plane, slab partitions keep the owned planes first).  0 when the mesh has no such structure."""
    grid = getattr(mesh, "grid", None)
    if grid is None:
        return 0
    nv = mesh.num_vertices()
    nplanes_v = None
    per_plane_v = int(np.prod([g + 1 for g in grid[:-1]]))
    if nv % per_plane_v != 0:
        return 0
    nplanes_v = nv // per_plane_v
    ndof = nv if vert2dof is None else int(np.max(vert2dof)) + 1
    if ndof % nplanes_v != 0:
        return 0
    stride = ndof // nplanes_v
    nown = ndof if n_owned is None else int(n_owned)
    return stride if nown % stride == 0 else 0


class TransportSolverF:
    """GPU state + operators of one mode-F problem.  Fields are torch fp64 tensors on the device."""

    def __init__(self, ctx: Context, mesh, model: TransportModel, bdata: BoundaryDataF, f_source=None,
                 rtol=1e-12, maxit=20000, check_every=0, halo=None, n_owned=None, c0_method="bicgstab", gmres_restart=30,
                 vert2dof=None, row_ordered=True, p_precond="line", cheb_degree=3, p_guess="previous", c0_precond="auto", fused_rows=False, shared_reaction_solve=True):
        """`vert2dof`: optional P1 dof of every vertex (periodic space); all fields, K, f_source and `bdata` are then in dof numbering."""
        self.ctx, self.model, self.rtol, self.maxit, self.check_every = ctx, model, rtol, maxit, check_every
        self.c0_method, self.gmres_restart = c0_method, gmres_restart
        # pressure PCG preconditioner: "jacobi" (default) or "chebyshev" (degree-k Chebyshev-Jacobi polynomial: same time on one GPU at
        # 200^3, 3-4x fewer reduction points per SpMV -- meant for multi-GPU runs; not yet the default there)
        self.p_method = {"chebyshev": "pcg_chebyshev", "line": "pcg_line"}.get(p_precond, "pcg")
        self.cheb_degree = cheb_degree
        self.fused_rows = bool(fused_rows)
        self.shared_reaction_solve = bool(shared_reaction_solve)
        self.p_guess = p_guess          # "extrapolate" (2 p^n - p^(n-1)) or "previous" (p^n) as the pressure PCG's initial iterate
        self._p_hist = 0                # pressure solutions of the current time series held in (p, p_old)
        self.halo = halo
        self.dm = DeviceMesh(ctx, mesh, vert2dof)
        V, Cn, d = self.dm.ndof, self.dm.ncell, self.dm.dim
        self.nloc = V
        self.nown = V if n_owned is None else int(n_owned)
        self.K_dev = None if np.isscalar(model.K) else ctx.to_device(np.asarray(model.K, dtype=np.float64))
        self.prm = transport_params(model.Pe, model.Da, model.phi, model.alpha, model.dt, model.sigma, model.rho0, model.gamma,
                                    model.fracs, model.zhat, model.K, model.reaction, model.ncomp, model.SUPG)
        # row dofmap: rows of ghost vertices (>= nown) are not assembled (owner-computes)
        rows = self.dm.cell_dofs
        if self.nown < V:
            rows = torch.where(rows < self.nown, rows, torch.full_like(rows, -1))
        self.plan = Plan(ctx, rows, self.dm.cell_dofs, self.nown, V)
        if row_ordered:
            self.dm.use_row_ordered_storage(self.plan)      # element kernels write in row order, the scatter streams
        l = d + 1
        self.emat = ctx.empty(Cn * l * l)
        self.evec = ctx.empty(Cn * l)
        self.evec2 = ctx.empty(Cn * l)
        # constant operators
        self.M = self.plan.new_matrix()
        check(ctx.lib.mp_elem_p1_mass(ctx.h, C.byref(self.dm.c), 1.0, ptr(self.emat)), "mp_elem_p1_mass")
        self.plan.scatter_matrix(self.emat, self.M)
        self.M.jacobi_setup()
        self.p_isbc = ctx.to_device(bdata.p_isbc, torch.uint8)
        self.c0_isbc = ctx.to_device(bdata.c0_isbc, torch.uint8)
        self.c0_g = ctx.to_device(bdata.c0_g)
        self.flux_load = ctx.to_device(bdata.flux_load)
        self.zeros_loc = ctx.zeros(V)
        self.L = self.plan.new_matrix()
        check(ctx.lib.mp_elem_p1_stiffness(ctx.h, C.byref(self.dm.c), ptr(self.K_dev), self.prm.Kconst, 1.0 / model.phi,
                                           ptr(self.emat)), "mp_elem_p1_stiffness")
        self.plan.scatter_matrix(self.emat, self.L)
        self.L.apply_dirichlet(None, self.p_isbc, self.zeros_loc)          # p = 0: homogeneous, matrix eliminated once
        self.L.jacobi_setup()
        # "line": block-tridiagonal Jacobi along the lattice lines of the LAST coordinate (rows are numbered plane by plane on a
        # BoxMesh / RectangleMesh, also inside a z-slab); falls back to point Jacobi on meshes without that structure
        self.line_stride = 0
        if self.p_method == "pcg_line":
            self.line_stride = lattice_line_stride(mesh, vert2dof, self.nown)
            if self.line_stride > 0 and not self._try(lambda: self.L.line_setup(self.line_stride)):
                self.line_stride = 0                 # a line block that is not SPD (should not happen for div(k grad p)): point Jacobi
            if self.line_stride == 0:
                self.p_method = "pcg"
        # the transport row uses the same lines (non-symmetric tridiagonal blocks, refactored with the matrix every step)
        self.c0_line = self.line_stride > 0 and c0_method == "bicgstab" and c0_precond != "jacobi"
        self.A = self.plan.new_matrix()
        # state
        self.u = ctx.zeros(Cn * d)
        self.p = ctx.zeros(V)
        self.p_old = ctx.zeros(V)
        self.c0 = ctx.zeros(V)
        self.c1 = ctx.zeros(V)
        self.c2 = ctx.zeros(V)
        self.c0_new = ctx.zeros(V)
        self.c1_new = ctx.zeros(V)
        self.c2_new = ctx.zeros(V)
        self.rhs = ctx.zeros(self.nown)
        self.rhs2 = ctx.zeros(self.nown)
        self.rhs_c0 = ctx.zeros(self.nown)
        self.fsrc = ctx.zeros(V) if f_source is None else ctx.to_device(np.asarray(f_source, dtype=np.float64))
        self.last = {}

    def set_state(self, c0=None, c1=None, c2=None, p=None, u=None):
        for name, v in (("c0", c0), ("c1", c1), ("c2", c2), ("p", p), ("u", u)):
            if v is not None:
                getattr(self, name).copy_(self.ctx.to_device(np.asarray(v, dtype=np.float64).ravel()))
                if name == "p":
                    self._p_hist = 0     # a new time series: no extrapolation across the jump

    def _exchange(self, x):
        if self.halo is not None:
            self.halo.exchange(x)

    def _solve_c0(self, kw):
        """Non-symmetric SUPG row: Jacobi-BiCGStab (2 SpMV + 3 fused vector kernels per iteration; 0.77 ms at 200^3 against
        1.85 ms for a GMRES(50) iteration) with restarted GMRES as the safeguard when BiCGStab breaks down or stagnates
        (it does on strongly non-normal rows, e.g. CFL ~ 400 with natural side walls).  The decision uses the allreduced
        residual, so every rank takes the same branch."""
        from ._lib import MpError
        if self.c0_method == "gmres":
            return self.A.solve("gmres", self.rhs_c0, self.c0_new, restart=self.gmres_restart, **kw)
        kb = dict(kw)
        kb["maxit"] = min(kw["maxit"], 1000)
        try:
            if self.c0_line and getattr(self, "_c0_line_ok", False):      # factors refreshed in assemble_c0_row
                r0 = self.A.solve("bicgstab_line", self.rhs_c0, self.c0_new, line_stride=self.line_stride, **kb)
            else:
                r0 = self.A.solve("bicgstab", self.rhs_c0, self.c0_new, **kb)
            if r0.converged:
                return r0
        except MpError as e:
            if "NaN" not in str(e):
                raise
        self.c0_new.copy_(self.c0)
        self.fallbacks = getattr(self, "fallbacks", 0) + 1
        return self.A.solve("gmres", self.rhs_c0, self.c0_new, restart=max(self.gmres_restart, 100), **kw)   # long recurrences: robust

    # ---- the pieces of one step (step() runs them in order; the at-size parity tests call them one by one)
    def assemble_reaction_rows(self):
        """rhs / rhs2 <- loads of the c1 (, c2) consistent-mass rows from the state (c0, c1, c2)."""
        ctx, lib, m, prm = self.ctx, self.ctx.lib, self.dm.c, self.prm
        three = self.model.ncomp == 3
        check(lib.mp_elem_p1_reaction_rhs(ctx.h, C.byref(m), C.byref(prm), ptr(self.c0), ptr(self.c1), ptr(self.c2) if three else None,
                                          ptr(self.evec), ptr(self.evec2) if three else None), "mp_elem_p1_reaction_rhs")
        self.plan.scatter_vector(self.evec, self.rhs)
        if three:
            self.plan.scatter_vector(self.evec2, self.rhs2)

    def _solve_reaction_rows_shared(self, kw):
        """The c1 and c2 rows (mupopp.py:936-937, 1234-1235) have the SAME load up to a factor:
             M c1' = M c1 - k2 T ,  M c2' = M c2 + k3 T ,   T_i = int w_i R(c0, c1),  k2/k3 = frac[1]/frac[2]
        so one mass solve for the common increment d = c1' - c1 = -k2 M^-1 T gives both: c1' = c1 + d, c2' = c2 - (frac[2]/frac[1]) d.
        Algebraically identical to the two solves of the reference's monolithic system; the tolerance applies to the increment, which
        is stricter for the fields themselves."""
        ctx, lib, m, prm = self.ctx, self.ctx.lib, self.dm.c, self.prm
        k2 = prm.dt * prm.frac[1] * prm.Da / (1.0 - prm.phi)
        check(lib.mp_elem_p1_reaction_load(ctx.h, C.byref(m), C.byref(prm), ptr(self.c0), ptr(self.c1), -k2, ptr(self.evec)), "mp_elem_p1_reaction_load")
        self.plan.scatter_vector(self.evec, self.rhs)                 # -k2 T, assembled directly (no cancellation against M c1)
        d = self.c2_new
        d.zero_()
        r1 = self.M.solve("pcg", self.rhs, d, **kw)
        self._exchange(d)
        self.c1_new.copy_(self.c1).add_(d)
        d.mul_(-float(self.model.fracs[2]) / float(self.model.fracs[1])).add_(self.c2)      # d is c2_new
        from .device import SolveResult
        return r1, SolveResult(0, r1.converged, r1.relres)

    def assemble_c0_row(self):
        """A, rhs <- Crank-Nicolson + SUPG row with the lagged velocity u^n and the NEW c1 (, c2); Dirichlet rows eliminated.
        `fused_rows=True` uses the fused row-wise kernel (mp_assemble_p1_transport_rows: CSR + SELL values, rhs, Dirichlet, Jacobi diagonal
        and the line-preconditioner bands in one pass, no element tensor: 8 GB of DRAM traffic instead of 18 GB at 200^3).  It is NOT the
        default: measured 10.5 ms against 10.2 ms for the element-kernel + scatter pipeline -- recomputing every cell once per incident
        row makes it issue/latency bound (4.5 G warp instructions, 24 % occupancy at 124 registers; profiles/r2_ncu_line_and_rows.md)."""
        ctx, lib, m, prm = self.ctx, self.ctx.lib, self.dm.c, self.prm
        three = self.model.ncomp == 3
        if self.fused_rows:
            A = self.A
            if A.dinv is None:
                A.dinv = ctx.empty(A.nrow)
            line = getattr(A, "_line", None)
            if self.c0_line and line is None:
                line = (self.line_stride,) + tuple(ctx.empty(A.nrow) for _ in range(3))
            st, lo, invw, mb = line if self.c0_line else (0, None, None, None)
            check(lib.mp_assemble_p1_transport_rows(ctx.h, C.byref(m), C.byref(prm), self.plan.h, ptr(self.u), ptr(self.c0), ptr(self.c1),
                                                    ptr(self.c2) if three else None, ptr(self.c1_new), ptr(self.c2_new) if three else None,
                                                    ptr(self.fsrc), ptr(self.c0_isbc), ptr(self.c0_g), ptr(A.vals),
                                                    ptr(A.sell_vals) if A.sell_vals is not None else None, ptr(self.rhs_c0), ptr(A.dinv),
                                                    int(st), ptr(lo), ptr(invw), ptr(mb)), "mp_assemble_p1_transport_rows")
            A.sell_dirty = A.sell_vals is None
            self._c0_line_ok = False
            if self.c0_line:
                self._c0_line_ok = self._try(lambda: check(lib.mp_line_factor(ctx.h, A.nrow, int(st), 0, ptr(lo), ptr(invw), ptr(mb)), "mp_line_factor"))
                A._line = (st, lo, invw, mb)
            return
        check(lib.mp_elem_p1_transport(ctx.h, C.byref(m), C.byref(prm), ptr(self.u), ptr(self.c0), ptr(self.c1),
                                       ptr(self.c2) if three else None, ptr(self.c1_new), ptr(self.c2_new) if three else None,
                                       ptr(self.fsrc), ptr(self.emat), ptr(self.evec)), "mp_elem_p1_transport")
        self.plan.scatter_matrix(self.emat, self.A)
        self.plan.scatter_vector(self.evec, self.rhs_c0)
        self.A.apply_dirichlet(self.rhs_c0, self.c0_isbc, self.c0_g)
        self.A.jacobi_setup()
        self._c0_line_ok = self.c0_line and self._try(lambda: self.A.line_setup(self.line_stride, spd=False))

    @staticmethod
    def _try(fn):
        """A line block of the transport row can have a zero pivot (no pivoting inside the tridiagonal LU): that step then uses point Jacobi.
        The factor kernel's verdict is a device flag read back by every rank, and the blocks are rank-local, so this is NOT collective --
        but the Jacobi and line variants issue the same exchanges per iteration, and convergence is decided on allreduced scalars."""
        from ._lib import MpError
        try:
            fn()
            return True
        except MpError as e:
            if "pivot" not in str(e) and "positive definite" not in str(e):
                raise
            return False

    def assemble_pressure_rhs(self):
        """rhs <- buoyancy load with the NEW c0 minus the prescribed boundary flux, p = 0 rows eliminated (the operator L is constant)."""
        ctx, lib, m, prm = self.ctx, self.ctx.lib, self.dm.c, self.prm
        check(lib.mp_elem_p1_pressure_rhs(ctx.h, C.byref(m), C.byref(prm), ptr(self.K_dev), ptr(self.c0_new), ptr(self.evec)),
              "mp_elem_p1_pressure_rhs")
        self.plan.scatter_vector(self.evec, self.rhs)
        ctx.axpby(-1.0, self.flux_load[: self.nown], 1.0, self.rhs)
        self.L.apply_dirichlet(self.rhs, self.p_isbc, self.zeros_loc, matrix=False)

    def recover_velocity(self):
        ctx, lib, m, prm = self.ctx, self.ctx.lib, self.dm.c, self.prm
        check(lib.mp_recover_velocity(ctx.h, C.byref(m), C.byref(prm), ptr(self.K_dev), ptr(self.p), ptr(self.c0_new), ptr(self.u)),
              "mp_recover_velocity")

    def phase_timing(self, on=True):
        """Record a CUDA event at every phase boundary of step() (on torch's current stream == the context's stream); phase_times()
        returns the accumulated milliseconds per phase.  Used by bench.py's untimed third pass, never inside a timed region."""
        self._pt = [] if on else None

    def phase_times(self):
        names = ("c1_c2_rows_and_solves", "c0_assembly", "c0_solve", "p_rhs", "p_solve", "velocity_recovery")
        out = dict.fromkeys(names, 0.0)
        torch.cuda.synchronize()
        for ev in self._pt or []:
            for k, nm in enumerate(names):
                out[nm] += ev[k].elapsed_time(ev[k + 1])
        return out

    def _mark(self, ev):
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append(e)

    def step(self):
        three = self.model.ncomp == 3
        ev = [] if getattr(self, "_pt", None) is not None else None
        self._mark(ev)
        kw = dict(rtol=self.rtol, maxit=self.maxit, halo=self.halo, check_every=self.check_every, refresh_jacobi=False)
        # 1. c1 (, c2): consistent-mass solves with the reaction load
        if three and self.shared_reaction_solve and self.model.fracs[1] != 0.0:
            r1, r2 = self._solve_reaction_rows_shared(kw)
        else:
            self.assemble_reaction_rows()
            self.c1_new.copy_(self.c1)
            r1 = self.M.solve("pcg", self.rhs, self.c1_new, **kw)
            self._exchange(self.c1_new)
            r2 = None
            if three:
                self.c2_new.copy_(self.c2)
                r2 = self.M.solve("pcg", self.rhs2, self.c2_new, **kw)
                self._exchange(self.c2_new)
        # 2. c0: CN + SUPG with the lagged velocity u^n
        self._mark(ev)
        self.assemble_c0_row()
        self._mark(ev)
        self.c0_new.copy_(self.c0)
        r0 = self._solve_c0(kw)
        self._exchange(self.c0_new)
        self._mark(ev)
        # 3. pressure with the NEW c0.  Initial guess: p^n, or the linear extrapolation 2 p^n - p^(n-1) once two solutions of THIS
        #    time series exist (fewer PCG iterations for the same answer: the tolerance is relative to ||b||, not to the guess)
        self.assemble_pressure_rhs()
        self._mark(ev)
        if self.p_guess == "extrapolate" and self._p_hist >= 2:
            self.p_old.mul_(-1.0).add_(self.p, alpha=2.0)            # p_old <- 2 p^n - p^(n-1): becomes the iterate
            self.p, self.p_old = self.p_old, self.p
        else:
            self.p_old.copy_(self.p)
        rp = self.L.solve(self.p_method, self.rhs, self.p, cheb_degree=self.cheb_degree, line_stride=self.line_stride, **kw)
        self._p_hist += 1
        self._exchange(self.p)
        self._mark(ev)
        # 4. velocity recovery, then rotate the state
        self.recover_velocity()
        self._mark(ev)
        if ev is not None:
            self._pt.append(ev)
        self.c0, self.c0_new = self.c0_new, self.c0
        self.c1, self.c1_new = self.c1_new, self.c1
        if three:
            self.c2, self.c2_new = self.c2_new, self.c2
        self.last = dict(c1=r1, c2=r2, c0=r0, p=rp)
        return self.last

    def true_residuals(self):
        """||b - A x|| / ||b|| of the LAST c0 and p solves, recomputed with one SpMV each (bench.py reports them; owned rows,
        allreduced over ranks).  Call right after step(): A / L / rhs still hold the p system, the c0 system is re-assembled."""
        ctx = self.ctx
        out = {}
        tmp = ctx.zeros(self.nown)

        def relres(A, b, x):
            self._exchange(x)
            A.spmv(x, tmp)
            ctx.axpby(1.0, b, -1.0, tmp)
            return float(np.sqrt(ctx.dot(tmp, tmp) / max(ctx.dot(b, b), 1e-300)))

        out["p"] = relres(self.L, self.rhs, self.p)
        out["c0"] = relres(self.A, self.rhs_c0, self.c0)          # the state was rotated: c0 is the solution of the last c0 row
        return out
