"""`core` helpers of the reference (/root/reference/src/modules/core.py) that sit on the per-timestep path: the CFL time step.

    u_max(U, cylinder_mesh)              core.py:23-39   max over cells of the cell average of |u|  (a DG0 projection + linf norm in DOLFIN)
    compute_dt(U, cfl, h_min, cylinder)  core.py:43-48   cfl * h_min / max(1e-6, u_max)
evaluated by one GPU kernel (`mp_max_cell_speed`); only the scalar comes back to the host.  `parse_param_file` / `int_float_string`
(core.py:52-115) are the same functions as in `mupopp_b200.mupopp`.
"""
from __future__ import annotations

import ctypes as C

from . import reference_element as ref
from ._lib import check
from .fenics import Function, _integrator, default_context
from .mupopp import int_float_string, parse_param_file  # noqa: F401

QDEG_SPEED = 4        # |u| is not polynomial; DOLFIN's automatic degree for sqrt(dot(u, u)) with P2 u is not reproducible offline


def u_max(U, cylinder_mesh=False):
    """Max over cells of (1/|cell|) int_cell |u| dx for the velocity part of U = (u, p[, o]) or a velocity Function."""
    u = U
    if not (isinstance(U, Function) and len(U.blocks) == U.space.mesh.dim):
        u = U.split()[0]
    ctx = default_context()
    mesh = u.space.mesh
    d = mesh.dim
    res = C.c_double()
    if u.cell_velocity is not None:                       # mode F: the cellwise velocity itself
        check(ctx.lib.mp_max_cell_speed(ctx.h, d, mesh.num_cells(), 2, C.c_void_p(u.cell_velocity.data_ptr()), None, None, None, 0, None, 0, None,
                                        C.byref(res)), "mp_max_cell_speed")
        return res.value
    deg = u.block_degree(0)
    it = _integrator(mesh)
    tab, w, ntab, nq, nb = it.tables("dx", QDEG_SPEED, deg)
    dofs = it.cell_dofs(deg, u.space.periodic.vert2dof if u.space.periodic is not None else None)
    p = [C.c_void_p(b.data_ptr()) for b in u.blocks] + [None] * (3 - d)
    check(ctx.lib.mp_max_cell_speed(ctx.h, d, mesh.num_cells(), 1, p[0], p[1], p[2], C.c_void_p(dofs.data_ptr()), nb, C.c_void_p(tab.data_ptr()), nq,
                                    C.c_void_p(w.data_ptr()), C.byref(res)), "mp_max_cell_speed")
    return res.value


def compute_dt(U, cfl, h_min, cylinder_mesh=False):
    """core.compute_dt: the CFL-limited time step."""
    return cfl * h_min / max(1.0e-6, u_max(U, cylinder_mesh))
