"""Bridge between the drop-in API objects (FunctionSpace, DirichletBC, Function of `mupopp_b200.fenics`) and the mode-F GPU timestep
(`transport.TransportSolverF`): what `solve(a == L, sol, bc)` runs for the multi-component forms when the space is periodic
(`constrained_domain`), the permeability is a P1 field, or mode F is requested.

Call sites served: /root/reference/src/run/CCS2D/CCS_3comp.py:206-263 (periodic CCS, constant K),
/root/reference/src/run/CCS2D/CCS_layered_permeability.py:211-254 (periodic, `K=darcy.perm` Expression, inflow `BoundarySource`),
/root/reference/src/run/CCS/CCS_3D_top_source.py:205-256.
Layout of a Function of the mixed space in mode F: [u_0..u_{d-1} (vertex averages, P1 dofs), p, c0, c1 (, c2)], all P1-sized; the
cellwise (DG0) velocity that the transport row actually uses travels as `Function.cell_velocity`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, ptr
from .transport import BoundaryDataF, TransportSolverF


def _value_signature(v):
    """Cheap fingerprint of a BC value / coefficient: changes when a Constant is re-assigned or an Expression attribute (t, dt, ...)
    is mutated, which is how the DOLFIN scripts make boundary data time dependent."""
    from .fenics import Constant, Expression, Function
    if isinstance(v, (int, float)):
        return ("num", float(v))
    if isinstance(v, Constant):
        return ("const", tuple(float(x) for x in v.value))
    if isinstance(v, Function):
        return ("fn", id(v), tuple(int(b._version) for b in v.blocks))
    if isinstance(v, Expression):
        items = []
        for k, x in sorted(vars(v).items()):
            if isinstance(x, (int, float, str, bool)):
                items.append((k, x))
        return ("expr", id(v), tuple(items))
    return ("obj", id(v))


def bc_signature(bcs):
    return tuple((id(bc), _value_signature(bc.value)) for bc in bcs)


def boundary_data_from_bcs(W, bcs):
    """transport.BoundaryDataF from DirichletBC objects: velocity BCs define Gamma_u and the flux load int (u_D.n) q ds, the rest of
    the (non-periodic) boundary gets p = 0; c0 BCs give the Dirichlet set of the transport row.  Host setup, vectorised where the
    marker / value allow it."""
    from .fenics import _evaluate
    mesh, per = W.mesh, W.periodic
    d = mesh.dim
    ub, c0b = list(W.sub_blocks[0]), W.sub_blocks[2][0]
    u_bcs, c0_bcs = [], []
    for bc in bcs:
        if bc.space is not W:
            raise ValueError("DirichletBC defined on a different FunctionSpace than the solution")
        if list(bc.block_ids) == ub:
            u_bcs.append(bc)
        elif list(bc.block_ids) == [c0b]:
            c0_bcs.append(bc)
        else:
            raise NotImplementedError("mode F takes Dirichlet conditions on the whole velocity (W.sub(0)) and on c0 (W.sub(2)); got blocks %r"
                                      % (bc.block_ids,))
    fv, nrm, meas, owner = mesh.boundary_facets()
    oloc = mesh.boundary_facet_local_index()
    if per is not None:                                          # periodic faces are not boundary
        keep = np.ones(fv.shape[0], dtype=bool)
        for a in per.axes:
            keep &= np.abs(nrm[:, a]) < 0.5
        fv, nrm, meas, owner, oloc = fv[keep], nrm[keep], meas[keep], owner[keep], oloc[keep]
    V = mesh.num_vertices()
    inside = np.zeros(fv.shape[0], dtype=bool)
    flux = np.zeros(V)
    mloc = (np.ones((d, d)) + np.eye(d)) / (d * (d + 1))
    for bc in u_bcs:
        sel = bc.facet_mask(fv, owner, oloc) & ~inside            # first BC wins on shared facets
        idx = np.nonzero(sel)[0]
        inside |= sel
        if idx.size == 0:
            continue
        pts = mesh.coords[fv[idx]].reshape(-1, d)
        cinfo = (np.repeat(owner[idx], d), np.repeat(oloc[idx], d))
        uD = _evaluate(bc.value, pts, d, cinfo).reshape(idx.size, d, d)           # (facet, vertex, component)
        g = np.einsum("fvc,fc->fv", uD, nrm[idx])
        np.add.at(flux, fv[idx].ravel(), (meas[idx][:, None] * (g @ mloc.T)).ravel())
    p_isbc = np.zeros(V, dtype=np.uint8)
    p_isbc[np.unique(fv[~inside].ravel())] = 1
    c0_isbc = np.zeros(V, dtype=np.uint8)
    c0_g = np.zeros(V)
    for bc in c0_bcs:
        dofs, vals = bc.dofs_and_values()[c0b]                   # vertex numbering (degree-1 block)
        c0_isbc[dofs] = 1
        c0_g[dofs] = vals
    if per is not None:
        g = np.zeros(per.ndof)
        m = c0_isbc.astype(bool)
        g[per.vert2dof[m]] = c0_g[m]
        return BoundaryDataF(per.fold_any(p_isbc), per.fold_sum(flux), per.fold_any(c0_isbc), g)
    return BoundaryDataF(p_isbc, flux, c0_isbc, c0_g)


class ModeFStepper:
    """One mode-F problem behind `solve(a == L, sol, bcs)`: owns the TransportSolverF, moves Function blocks in and out."""

    def __init__(self, ctx, W, bcs, model, f_nodal, rtol=1e-12, maxit=50000, p_precond="line", lagged_velocity="cache"):
        self.ctx, self.W, self.bcs = ctx, W, list(bcs)       # references keep the ids used in cache keys alive
        per = W.periodic
        self.sig = bc_signature(bcs)
        bd = boundary_data_from_bcs(W, bcs)
        self.s = TransportSolverF(ctx, W.mesh, model, bd, f_source=f_nodal, rtol=rtol, maxit=maxit,
                                  vert2dof=per.vert2dof if per is not None else None, p_guess="previous", p_precond=p_precond)
        self._lump = None
        self._edges = None
        self._tmp = ctx.zeros(self.s.nown)
        # "cache": the lagged cellwise velocity u^n travels with the previous Function (Function.cell_velocity, device resident);
        # "recover": u^n is recomputed from (p^n, c0^n) at the start of the solve -- mode F DEFINES u as that function of the state
        # (TransportSolverF.recover_velocity), so a host-resident workflow need not ship the 3*ncell doubles of the DG0 field.
        # The two agree bit for bit whenever the previous state came out of a mode-F solve; they differ only for a user-supplied
        # initial velocity (ignored by "recover").
        if lagged_velocity not in ("cache", "recover"):
            raise ValueError("mupopp_b200_lagged_velocity must be 'cache' or 'recover'")
        self.lagged_velocity = lagged_velocity

    def refresh(self, bcs, model, f_nodal):
        """Per solve: new physical parameters (dt may change), boundary VALUES if their fingerprint changed, source term."""
        from .device import transport_params
        s = self.s
        sig = bc_signature(bcs)
        if sig != self.sig:
            bd = boundary_data_from_bcs(self.W, bcs)
            if not (np.array_equal(bd.p_isbc, s.p_isbc.cpu().numpy()) and np.array_equal(bd.c0_isbc, s.c0_isbc.cpu().numpy())):
                raise NotImplementedError("the Dirichlet SETS changed between solves; create a new FunctionSpace/solver for a new boundary layout")
            s.c0_g.copy_(self.ctx.to_device(bd.c0_g))
            s.flux_load.copy_(self.ctx.to_device(bd.flux_load))
            self.bcs, self.sig = list(bcs), sig
        s.model = model
        s.prm = transport_params(model.Pe, model.Da, model.phi, model.alpha, model.dt, model.sigma, model.rho0, model.gamma,
                                 model.fracs, model.zhat, model.K, model.reaction, model.ncomp, model.SUPG)
        if f_nodal is not None:
            s.fsrc.copy_(self.ctx.to_device(np.asarray(f_nodal, dtype=np.float64)))

    def load_state(self, prev):
        s, d = self.s, self.W.mesh.dim
        nc = s.model.ncomp
        if self.lagged_velocity == "recover":
            s.p.copy_(prev.blocks[d])
            s.c0_new.copy_(prev.blocks[d + 1])
            s.recover_velocity()                     # u^n = -(K/phi) (grad p^n + sigma rho(c0^n) zhat), cell by cell
        elif getattr(prev, "cell_velocity", None) is not None:
            s.u.copy_(prev.cell_velocity)
        else:      # a user-supplied nodal velocity (the scripts start from zero): cell value = mean of the cell's vertex values
            cd = s.dm.cell_dofs.long()       # (P2 blocks of a non-periodic space: their first entries are the vertex values)
            s.u.copy_(torch.stack([prev.blocks[c][cd].mean(1) for c in range(d)], 1).reshape(-1))
        s.p.copy_(prev.blocks[d])
        s.c0.copy_(prev.blocks[d + 1])
        s.c1.copy_(prev.blocks[d + 2])
        if nc == 3:
            s.c2.copy_(prev.blocks[d + 3])

    def store_state(self, out):
        s, d, ctx = self.s, self.W.mesh.dim, self.ctx
        nc = s.model.ncomp
        out.blocks[d].copy_(s.p)
        out.blocks[d + 1].copy_(s.c0)
        out.blocks[d + 2].copy_(s.c1)
        if nc == 3:
            out.blocks[d + 3].copy_(s.c2)
        if out.cell_velocity is None:
            out.cell_velocity = s.u.clone()
        else:
            out.cell_velocity.copy_(s.u)
        # nodal velocity for output: volume-weighted vertex average of the cellwise recovery
        m = s.dm.c
        if self._lump is None:
            check(ctx.lib.mp_elem_p1_cell_load(ctx.h, C.byref(m), None, 0, ptr(s.evec)), "mp_elem_p1_cell_load")
            self._lump = ctx.zeros(s.nown)
            s.plan.scatter_vector(s.evec, self._lump)
            self._lump.reciprocal_()
        for c in range(d):
            check(ctx.lib.mp_elem_p1_cell_load(ctx.h, C.byref(m), ptr(s.u), c, ptr(s.evec)), "mp_elem_p1_cell_load")
            s.plan.scatter_vector(s.evec, self._tmp)
            blk = out.blocks[c]
            if blk.numel() == s.nown:
                ctx.vmul(1.0, self._lump, self._tmp, blk)
            else:      # a non-periodic space keeps its P2 velocity blocks: vertex values + edge-midpoint means (the P1 field written in P2)
                ctx.vmul(1.0, self._lump, self._tmp, self._tmp)
                if self._edges is None:
                    self._edges = ctx.to_device(self.W.mesh.edges.astype(np.int64), torch.int64)
                blk[: s.nown].copy_(self._tmp)
                blk[s.nown:].copy_(0.5 * (self._tmp[self._edges[:, 0]] + self._tmp[self._edges[:, 1]]))

    def step(self, prev, out):
        self.load_state(prev)
        res = self.s.step()
        for name, r in res.items():
            # `converged` refers to the TRUE residual, re-evaluated after the recurrences stop.  On badly conditioned operators (e.g. a
            # 1000x permeability contrast) fp64 recurrences stagnate a little above a 1e-12 tolerance: such a solve is as converged as
            # the reference's sparse LU (whose residual is ~ eps * cond as well) and is accepted; anything worse is an error.
            if r is not None and not r.converged and not (r.relres <= 1e3 * self.s.rtol):
                raise RuntimeError("solve: the %s Krylov solve did not converge (%r)" % (name, r))
        self.store_state(out)
        return res
