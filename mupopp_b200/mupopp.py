"""The `mupopp` module API on the B200 path: same classes, method names, positional order and defaults as
/root/reference/src/modules/mupopp.py (DarcyAdvection :560, CCS :1104, StokesAdvection :965, compaction :125, utilities :23-115).

As in the reference, the form methods do not solve anything: they return `(a, L)` descriptors; `solve(a == L, sol, bcs)` from
`mupopp_b200.fenics` runs them.  The solve is the reference's own discretisation (mode R: mixed P2/P1, segregated
block-triangular sequence == the monolithic LU solve) executed by the CUDA kernels of libmupopp_b200.so.
Forms whose kernels are not built yet raise NotImplementedError at solve time with the row of SURVEY.md section 8 they belong to.
"""
from __future__ import annotations

import numpy as np

from .fenics import Constant, DirichletBC, Expression, Form, Function, default_context, info  # noqa: F401


# ------------------------------------------------------------------ utilities (mupopp.py:23-115)
def symgrad(u):
    raise NotImplementedError("symgrad builds a UFL expression; UFL is not part of the B200 path")


def iteration_output_info(niter, h, l2_u, l2_p, l2_c):
    print("Number of Krylov iterations:", niter)
    print("hmax:", h)
    print("L2 error in velocity:", l2_u)
    print("L2 error in pressure:", l2_p)
    print("L2 error in compaction:", l2_c)


def iteration_output_write(niter, h, l2_u, l2_p, l2_c, fname="output/iter.out"):
    import os
    os.makedirs(os.path.dirname(fname) or ".", exist_ok=True)
    with open(fname, "a") as f:
        f.write("%d,%g,%g,%g,%g\n" % (niter, h, l2_u, l2_p, l2_c))


def int_float_string(val):
    if val.isdigit():
        return int(val)
    try:
        return float(val)
    except ValueError:
        return val.replace(" ", "")


def parse_param_file(filename):
    """`key = value  # comment` files (e.g. benchmark/time_marching/time_test1.cfg) -> dict with int/float/str values."""
    out = {}
    with open(filename, "r") as f:
        for raw in f.read().splitlines():
            if raw.strip() == "" or raw.startswith("#"):
                continue
            opt = raw.rstrip().replace(" ", "").partition("#")[0]
            if "=" not in opt:
                continue
            key, val = opt.split("=", 1)
            out[key] = int_float_string(val)
    return out


def _zhat(zh, dim):
    v = np.asarray(getattr(zh, "value", zh), dtype=float).ravel()
    return tuple(float(v[i]) if i < v.size else 0.0 for i in range(dim))


def _kvalue(K):
    if isinstance(K, (int, float)):
        return float(K)
    if isinstance(K, Constant):
        return float(K)
    raise NotImplementedError("mode R on the GPU supports a constant permeability K (P1 K: mode F, mupopp_b200.transport)")


# ---- which discretisation runs a multi-component solve
#   "R": the reference's mixed P2/P1 space, segregated == monolithic LU (parity mode; constant K, no periodic faces yet)
#   "F": primal P1 pressure operator + cellwise velocity recovery (BASELINE.json north_star (1)-(3); the benchmarked path)
#   "auto" (default): F when the FunctionSpace has a periodic constrained_domain or K is a P1 field, else R.
# Set with mupopp_b200.mupopp.set_mode(...), the environment variable MUPOPP_B200_MODE, or per call with
# solve(..., solver_parameters={"mupopp_b200_mode": "F"}).
import os as _os

_MODE = _os.environ.get("MUPOPP_B200_MODE", "auto").upper()


def set_mode(mode):
    global _MODE
    mode = str(mode).upper()
    if mode not in ("R", "F", "AUTO"):
        raise ValueError("mode must be 'R', 'F' or 'auto'")
    _MODE = mode


def _k_field(K, W):
    """Permeability as the forms take it: a float, or the nodal values of a P1 field (Expression / Function, e.g. `K=darcy.perm`,
    run/CCS2D/CCS_layered_permeability.py:233,250) at the P1 dofs of W.  Evaluated once per (coefficient, attribute fingerprint)."""
    if isinstance(K, (int, float, Constant)):
        return float(K)
    from .modef_api import _value_signature
    cache = W.__dict__.setdefault("_coef_cache", {})
    key = ("K", _value_signature(K))
    if key not in cache:
        if isinstance(K, Function):
            cache[key] = (K, K.blocks[0].cpu().numpy().copy())
        else:
            from .fenics import _evaluate
            cache[key] = (K, _evaluate(K, W.dof_coords(1), 1)[:, 0].copy())
    return cache[key][1]


def _pair(kind, **data):
    return Form(kind, "a", **data), Form(kind, "L", **data)


# ------------------------------------------------------------------ classes
class DarcyAdvection:
    """mupopp.py:560-960."""

    def __init__(self, Pe=100, Da=10.0, phi=0.01, alpha=0.005, cfl=1.0e-2, dt=1.0e-2):
        self.Pe, self.Da, self.phi, self.cfl, self.dt, self.alpha = Pe, Da, phi, cfl, dt, alpha
        self.beta = alpha / phi

    def darcy_bilinear(self, W, mesh, K=0.1, zh=Constant((0.0, 1.0)), TwodTrue=True):
        zhat = _zhat(zh, 2) if TwodTrue else (0.0, 0.0, 1.0)
        return _pair("darcy_bilinear", W=W, mesh=mesh, K=_kvalue(K), zhat=zhat)

    def advection_diffusion(self, Q, u0, velocity, dt, mesh):
        return _pair("advection_diffusion", Q=Q, u0=u0, velocity=velocity, dt=float(dt), mesh=mesh, Pe=float(self.Pe), Da=float(self.Da))

    def advection_diffusion_two_component_adaptive(self, Q, c_prev, velocity, mesh):
        return _pair("two_component_adaptive", Q=Q, c_prev=c_prev, velocity=velocity, mesh=mesh, owner=self)

    def _multi(self, kind, W, mesh, sol_prev, dt, f1, K, zh, SUPG, gam_rho, rho_0):
        from .transport import TransportModel
        mk = TransportModel.darcy_two_component if kind == 2 else TransportModel.darcy_three_component
        model = mk(Pe=self.Pe, Da=self.Da, phi=self.phi, alpha=self.alpha, dt=float(getattr(dt, "dt", dt)), K=_k_field(K, W),
                   zh=_zhat(zh, mesh.dim), SUPG=SUPG, gam_rho=gam_rho, rho_0=rho_0)
        return _pair("multi_component", W=W, mesh=mesh, sol_prev=sol_prev, f=f1, model=model)

    def advection_diffusion_two_component(self, W, mesh, sol_prev, dt, f1, K=1.0, zh=Constant((0.0, 1.0)), SUPG=1, gam_rho=-1.0, rho_0=1.0):
        return self._multi(2, W, mesh, sol_prev, dt, f1, K, zh, SUPG, gam_rho, rho_0)

    def advection_diffusion_three_component(self, W, mesh, sol_prev, dt, f1, K=1.0, zh=Constant((0.0, 1.0)), SUPG=1, gam_rho=-1.0, rho_0=1.0):
        return self._multi(3, W, mesh, sol_prev, dt, f1, K, zh, SUPG, gam_rho, rho_0)


class CCS:
    """mupopp.py:1104-1262."""

    def __init__(self, Pe=100, Da=10.0, phi=0.01, alpha=0.005, cfl=1.0e-2, dt=1.0e-2):
        self.Pe, self.Da, self.phi, self.cfl, self.dt, self.alpha = Pe, Da, phi, cfl, dt, alpha
        self.beta = alpha / phi

    def advection_diffusion_three_component(self, W, mesh, sol_prev, dt, f_source, K=1.0, zh=Constant((0.0, 1.0)), SUPG=1, gam_rho=1.0):
        from .transport import TransportModel
        model = TransportModel.ccs_three_component(Pe=self.Pe, Da=self.Da, phi=self.phi, alpha=self.alpha, dt=float(getattr(dt, "dt", dt)),
                                                   K=_k_field(K, W), zh=_zhat(zh, mesh.dim), SUPG=SUPG, gam_rho=gam_rho)
        return _pair("multi_component", W=W, mesh=mesh, sol_prev=sol_prev, f=f_source, model=model)


class StokesAdvection:
    """mupopp.py:965-1099 (P3 velocity Stokes + ADR + precipitation), SURVEY.md section 8 row a7."""

    def __init__(self, Pe=100, Da=10.0, alpha=0.005, cfl=1.0e-2, dt=1.0e-2):
        self.Pe, self.Da, self.cfl, self.dt, self.alpha = Pe, Da, cfl, dt, alpha

    def stokes_ADR_precipitation(self, W, mesh, sol_prev, dt, SUPG=2):
        return _pair("stokes_adr", W=W, mesh=mesh, sol_prev=sol_prev, dt=dt, SUPG=SUPG, owner=self)


class compaction:
    """mupopp.py:125-471: parameters come from a `key = value` file; forms are rows a8-a10 of SURVEY.md section 8."""

    def __init__(self, param_file, ksp="minres"):
        param = parse_param_file(param_file)
        self.da, self.R, self.B = param["da"], param["R"], param["B"]
        self.theta, self.dL, self.T = param["theta"], param["dL"], param["T"]
        self.dt, self.out_freq = param["dt"], param["out_freq"]
        self.krylov_method = ksp

    def surface_tension_2(self, phi):
        # mupopp.py:206-215: chi(phi) polynomial, theta passed raw to cos (radians)
        p1, p2, p3, p4 = -8065.0, 6149.0, -1778.0, 249.0
        return (2.0 * np.cos(self.theta) - 1.0) * (2.0 * p4 + 20.0 * p1 * phi ** 3 + 12.0 * p2 * phi ** 2 + 6.0 * p3 * phi)

    def mass_conservation(self, V, phi0, u, dt, gam, mesh):
        a, L = _pair("compaction_mass", V=V, phi0=phi0, u=u, dt=dt, gam=gam, mesh=mesh, owner=self)
        return a, L, Form("compaction_mass", "b", **a.data)

    def mass_solver(self, X, a_phi, L_phi, bb_phi, nits=100, tol=1.0e-6, monitor=False):
        """mupopp.py:261-304: assemble the porosity row and solve it (Jacobi-GMRES on the GPU; the reference runs `krylov_method`
        with AMG on this non-symmetric matrix).  Returns the new porosity Function."""
        sol = Function(X)
        _solve_compaction_mass(a_phi, sol, tol, max(nits, 1), monitor)
        return sol

    def momentum_conservation(self, W, phi, gam, buyoancy, zh=Constant((0.0, 0.0, 1.0))):
        a, L = _pair("compaction_momentum", W=W, phi=phi, gam=gam, buoyancy=buyoancy, zh=zh, owner=self)
        return a, L, Form("compaction_momentum", "b", **a.data)

    def momentum_solver(self, W, a, L, b, bcs, pc="amg", tol=1.0e-6, max_its=3000, monitor=False):
        """mupopp.py:383-429: assemble_system(a, L, bcs), assemble_system(b, L, bcs), MINRES -> (u, p, c, niter).
        On the GPU the preconditioner is built from the reference's own form `b` (which the reference hands to hypre AMG)."""
        bcs = list(bcs) if isinstance(bcs, (list, tuple)) else [bcs]
        sol = Function(W)
        niter = _solve_compaction_momentum(a, sol, bcs, tol, max_its, monitor)
        u, p, c = sol.split()
        return u, p, c, niter

    def momentum_solver_direct(self, W, a, L, b, bcs, pc="amg", tol=1.0e-6, max_its=3000, monitor=False):
        """mupopp.py:430-471 solves the same system with LU (quadrature_degree=3); here: the same MINRES driven to a tight tolerance."""
        return self.momentum_solver(W, a, L, b, bcs, pc=pc, tol=1.0e-11, max_its=max(max_its, 50000), monitor=monitor)


# ------------------------------------------------------------------ solve dispatch (called by fenics.solve)
_CACHE = {}


def _solve_compaction_mass(form, sol, tol, maxit, monitor=False):
    """compaction.mass_conservation row (mupopp.py:217-259) -> sol; shared by compaction.mass_solver and the KrylovSolver shim."""
    from . import compaction as _c, moder
    d = form.data
    own = d["owner"]
    deg = d["V"].blocks[0][0]
    if deg not in (1, 2):
        raise NotImplementedError("compaction.mass_solver: the porosity space must be CG1 or CG2")
    if d["phi0"].block_degree(0) != deg or d["gam"].block_degree(0) != deg:
        raise NotImplementedError("compaction.mass_solver: phi0 and gam must live in the porosity space X (as in the reference's scripts)")
    ctx = default_context()
    mesh = d["mesh"]
    sp = _spaces(ctx, mesh)
    key = ("porosity", id(mesh), float(own.da), deg)
    if key not in _CACHE:
        _CACHE[key] = _c.PorosityR(ctx, sp, own.da, degree=deg)
    dt = float(getattr(d["dt"], "dt", d["dt"]))
    res = _CACHE[key].solve(d["phi0"].blocks[0].clone(), d["u"].blocks, dt, d["gam"].blocks[0], sol.blocks[0], rtol=tol, maxit=maxit)
    if monitor:
        info("mass_solver: %s" % (res,))
    return res.niter


def _solve_compaction_momentum(form, sol, bcs, tol, maxit, monitor=False):
    """compaction.momentum_conservation system (mupopp.py:306-382) -> sol = [u, p, c]; returns the MINRES iteration count."""
    from . import compaction as _c
    d = form.data
    own = d["owner"]
    W = d["W"]
    ctx = default_context()
    mesh = W.mesh
    dim = mesh.dim
    if d["phi"].space.blocks[0][0] != 1:
        raise NotImplementedError("compaction.momentum_solver: only P1 phi/gam/buoyancy coefficients are on the B200 path")
    sp = _spaces(ctx, mesh)
    key = ("compaction", id(W), _bc_sig(bcs), own.da, own.R, own.B, own.theta, own.dL, tuple(_zhat(d["zh"], dim)))
    if key not in _CACHE:
        ub = W.sub_blocks[0]
        arr = _bc_arrays(W, bcs, list(ub))
        s_ = _c.CompactionMomentumR(ctx, sp, own.da, own.R, own.B, own.theta, own.dL, _zhat(d["zh"], dim),
                                    [arr[k][0] for k in ub], [arr[k][1] for k in ub])
        s_._keep = (W, list(bcs))
        _CACHE[key] = s_
    s = _CACHE[key]
    s.assemble(d["phi"].blocks[0], d["gam"].blocks[0], d["buoyancy"].blocks[0])
    x = _c.BlockVector(ctx, dim, sp.n2, sp.n1)
    niter = s.solve(x, rtol=tol, maxit=maxit)
    if monitor:
        info("momentum_solver: %d MINRES iterations, relative residual %.3e" % (niter, s.last_relres))
    for k in range(dim):
        sol.blocks[k].copy_(x.u[k])
    sol.blocks[dim].copy_(x.p)
    sol.blocks[dim + 1].copy_(x.c)
    return niter


def _spaces(ctx, mesh):
    from . import moder
    key = ("spaces", id(mesh))
    if key not in _CACHE:
        sp = moder.Spaces(ctx, mesh)
        sp._keep = mesh
        _CACHE[key] = sp
    return _CACHE[key]


def _krylov_dispatch(form, sol, bcs, tol, maxit, monitor):
    """KrylovSolver(...).solve(U.vector(), b) / LUSolver on operators from assemble_system (run/sphere_archer/sphere.py:193-207,256-286)."""
    if form.kind == "compaction_momentum":
        return _solve_compaction_momentum(form, sol, bcs, tol, maxit, monitor)
    if form.kind == "compaction_mass":
        return _solve_compaction_mass(form, sol, tol, maxit, monitor)
    return _dispatch_solve(form, sol, bcs, {"relative_tolerance": tol})


def _bc_arrays(space, bcs, block_ids):
    """isbc / g arrays per requested scalar block from a list of DirichletBC (later BCs override earlier ones)."""
    out = {}
    for b in block_ids:
        nd = space.blocks[b][2]
        out[b] = (np.zeros(nd, dtype=np.uint8), np.zeros(nd))
    for bc in bcs:
        if bc.space is not space:
            raise ValueError("DirichletBC defined on a different FunctionSpace than the solution")
        for b, (dofs, vals) in bc.dofs_and_values().items():
            if b in out and dofs.size:
                out[b][0][dofs] = 1
                out[b][1][dofs] = vals
    return out


def _bc_key(bcs):
    return tuple(id(b) for b in bcs)


def _bc_sig(bcs):
    from .modef_api import bc_signature
    return bc_signature(bcs)


def _source_nodal(f, W):
    """Source term (user Expression subclass, Function, Constant, number) at the P1 dofs of W; evaluated once per fingerprint -- the
    run scripts rebuild the forms every timestep with the same source object (run/CCS2D/CCS_3comp.py:249,260)."""
    if f is None:
        return None
    from .modef_api import _value_signature
    cache = W.__dict__.setdefault("_coef_cache", {})
    key = ("f", _value_signature(f))
    if key not in cache:
        if isinstance(f, Function):
            cache[key] = (f, f.blocks[0].cpu().numpy().copy())
        else:
            from .fenics import _evaluate
            cache[key] = (f, _evaluate(f, W.dof_coords(1), 1)[:, 0].copy())
    return cache[key][1]


def _nodal(f, space_mesh):
    """A P1 coefficient given as Function / Expression / Constant / number -> nodal numpy array."""
    from .fenics import _evaluate
    if f is None:
        return None
    if isinstance(f, Function):
        return f.blocks[0].cpu().numpy()
    return _evaluate(f, space_mesh.coords, 1)[:, 0]


def _dispatch_solve(form, u, bcs, solver_parameters):
    from . import moder
    ctx = default_context()
    d = form.data
    kind = form.kind
    if kind == "darcy_bilinear":
        W, mesh = d["W"], d["mesh"]
        dim = mesh.dim
        key = ("darcy", id(W), d["K"], _bc_sig(bcs))
        if key not in _CACHE:
            sp = _spaces(ctx, mesh)
            arr = _bc_arrays(W, bcs, W.sub_blocks[0])
            isbc = [arr[b][0] for b in W.sub_blocks[0]]
            g = [arr[b][1] for b in W.sub_blocks[0]]
            _CACHE[key] = moder.MixedDarcyR(ctx, sp, K=d["K"], phi=1.0, u_isbc=isbc, u_g=g)
            _CACHE[key]._keep = (W, list(bcs))
        darcy = _CACHE[key]
        sp = darcy.sp
        import torch
        ones = torch.ones(sp.n1, dtype=torch.float64, device=ctx.device)
        lK = ctx.zeros(sp.n2)
        darcy.G.spmv(ones, lK)                                   # L = K f v.zhat with f == 1 (mupopp.py:616,621)
        rhs = []
        for c in range(dim):
            rc = ctx.zeros(sp.n2)
            ctx.axpby(float(d["zhat"][c]), lK, 0.0, rc)
            rhs.append(rc)
        p = u.blocks[dim]
        uu, its = darcy.solve(rhs, p)
        for c in range(dim):
            u.blocks[c].copy_(uu[c])
        return its
    if kind == "advection_diffusion":
        Q, mesh = d["Q"], d["mesh"]
        key = ("adr", id(Q), d["Pe"], d["Da"], d["dt"], _bc_key(bcs))
        if key not in _CACHE:
            sp = _spaces(ctx, mesh)
            arr = _bc_arrays(Q, bcs, [0])
            vdeg = d["velocity"].block_degree(0)
            _CACHE[key] = moder.AdrSolverR(ctx, sp, moder.adr_single_params(d["Pe"], d["Da"], d["dt"]), arr[0][0], arr[0][1], vel_degree=vdeg)
            _CACHE[key]._keep, _CACHE[key]._bc_sig = (Q, list(bcs)), _bc_sig(bcs)
        adr = _CACHE[key]
        if _bc_sig(bcs) != adr._bc_sig:              # time-dependent boundary values (Constant.assign / Expression attribute)
            adr.g.copy_(ctx.to_device(np.asarray(_bc_arrays(Q, bcs, [0])[0][1], dtype=np.float64)))
            adr._bc_sig = _bc_sig(bcs)
        c_in = d["u0"].blocks[0].clone()             # u0 may alias the solution Function (c0 = sol_c in the script)
        res = adr.solve(d["velocity"].blocks, c_in, u.blocks[0])
        return res.niter
    if kind == "multi_component":
        W, mesh, model = d["W"], d["mesh"], d["model"]
        dim = mesh.dim
        nc = model.ncomp
        mode = str(solver_parameters.get("mupopp_b200_mode", _MODE)).upper()
        if mode == "AUTO":
            mode = "F" if (W.periodic is not None or not np.isscalar(model.K)) else "R"
        fs = _source_nodal(d["f"], W)
        if mode == "F":
            from .modef_api import ModeFStepper
            # "preconditioner" (dolfin's solver_parameters key): "jacobi" | "chebyshev" | "line" (lattice-line block Jacobi on structured
            # meshes, point Jacobi elsewhere) | "default" (= "line")
            pc = str(solver_parameters.get("preconditioner", "default")).lower()
            pc = pc if pc in ("jacobi", "chebyshev", "line") else "line"        # DOLFIN names ("amg", "ilu", "default", ...) -> the best one here
            lag = str(solver_parameters.get("mupopp_b200_lagged_velocity", "cache")).lower()
            key = ("multiF", id(W), _bc_key(bcs), model.reaction, nc, model.SUPG, model.sigma, id(model.K) if not np.isscalar(model.K) else model.K, pc, lag)
            if key not in _CACHE:
                _CACHE[key] = ModeFStepper(ctx, W, bcs, model, fs, rtol=float(solver_parameters.get("relative_tolerance", 1e-12)), p_precond=pc,
                                           lagged_velocity=lag)
            st = _CACHE[key]
            st.refresh(bcs, model, fs)
            return st.step(d["sol_prev"], u)
        if W.periodic is not None:
            raise NotImplementedError("mode R has no periodic constrained_domain yet (SURVEY.md section 8 f1); use mode F")
        if not np.isscalar(model.K):
            raise NotImplementedError("mode R on the GPU supports a constant permeability K; a P1 K runs in mode F")
        key = ("multi", id(W), _bc_key(bcs), model.reaction, nc, model.SUPG, model.K, model.phi)
        if key not in _CACHE:
            ub = W.sub_blocks[0]
            arr = _bc_arrays(W, bcs, list(ub) + [W.sub_blocks[2][0]])
            s_ = moder.TransportSolverR(ctx, mesh, model, [arr[b][0] for b in ub], [arr[b][1] for b in ub],
                                        arr[W.sub_blocks[2][0]][0], arr[W.sub_blocks[2][0]][1], f_source=fs)
            s_._keep = (W, list(bcs))           # keeps the ids in the cache key valid
            s_._bc_sig = _bc_sig(bcs)
            _CACHE[key] = s_
        s = _CACHE[key]
        if _bc_sig(bcs) != s._bc_sig:           # boundary VALUES are re-evaluated when a Constant / Expression attribute changed
            ub = W.sub_blocks[0]
            arr = _bc_arrays(W, bcs, list(ub) + [W.sub_blocks[2][0]])
            for c, b in enumerate(ub):
                s.darcy.g[c].copy_(ctx.to_device(np.asarray(arr[b][1], dtype=np.float64)))
            s.adr.g.copy_(ctx.to_device(np.asarray(arr[W.sub_blocks[2][0]][1], dtype=np.float64)))
            s._bc_sig = _bc_sig(bcs)
        s.model, s.prm = model, moder.transport_params(model.Pe, model.Da, model.phi, model.alpha, model.dt, model.sigma, model.rho0,
                                                        model.gamma, model.fracs, model.zhat, model.K, model.reaction, model.ncomp, model.SUPG)
        s.adr.prm = s.prm
        prev = d["sol_prev"]
        for c in range(dim):
            s.u[c].copy_(prev.blocks[c])
        s.p.copy_(prev.blocks[dim])
        s.c0.copy_(prev.blocks[dim + 1])
        s.c1.copy_(prev.blocks[dim + 2])
        if nc == 3:
            s.c2.copy_(prev.blocks[dim + 3])
        if fs is not None:
            s.fsrc.copy_(ctx.to_device(np.asarray(fs, dtype=np.float64)))
        info_ = s.step()
        for c in range(dim):
            u.blocks[c].copy_(s.u[c])
        u.blocks[dim].copy_(s.p)
        u.blocks[dim + 1].copy_(s.c0)
        u.blocks[dim + 2].copy_(s.c1)
        if nc == 3:
            u.blocks[dim + 3].copy_(s.c2)
        return info_
    if kind == "two_component_adaptive":
        Q, mesh, own = d["Q"], d["mesh"], d["owner"]
        key = ("adaptive", id(Q), own.Pe, own.Da, own.phi, own.alpha, own.dt, _bc_sig(bcs))
        if key not in _CACHE:
            sp = _spaces(ctx, mesh)
            arr = _bc_arrays(Q, bcs, [Q.sub_blocks[0][0]])
            b0 = Q.sub_blocks[0][0]
            _CACHE[key] = moder.TwoComponentAdaptiveR(ctx, sp, own.Pe, own.Da, own.phi, own.alpha, own.dt, arr[b0][0], arr[b0][1],
                                                      vel_degree=d["velocity"].block_degree(0))
            _CACHE[key]._keep = (Q, list(bcs))
        s = _CACHE[key]
        prev = d["c_prev"]
        c0n, c1n = prev.blocks[0].clone(), prev.blocks[1].clone()
        res = s.step(d["velocity"].blocks, c0n, c1n, u.blocks[0], u.blocks[1])
        return res.niter
    if kind == "stokes_adr":
        from . import stokes as _st
        W, mesh, own = d["W"], d["mesh"], d["owner"]
        dim = mesh.dim
        vdeg = W.blocks[W.sub_blocks[0][0]][0]
        dt = float(getattr(d["dt"], "dt", d["dt"]))
        key = ("stokes", id(W), _bc_sig(bcs), own.Pe, own.Da, dt, d["SUPG"])
        if key not in _CACHE:
            ub = W.sub_blocks[0]
            c0b = W.sub_blocks[2][0]
            arr = _bc_arrays(W, bcs, list(ub) + [c0b])
            _CACHE[key] = _st.StokesAdrR(ctx, mesh, own.Pe, own.Da, dt, d["SUPG"], [arr[k][0] for k in ub], [arr[k][1] for k in ub],
                                         arr[c0b][0], arr[c0b][1], vdeg=vdeg)
            _CACHE[key]._keep = (W, list(bcs))
        s = _CACHE[key]
        prev = d["sol_prev"]
        uv = s.u_views()
        for c in range(dim):
            uv[c].copy_(prev.blocks[c])
        s.p.copy_(prev.blocks[dim])
        s.c0.copy_(prev.blocks[dim + 1])
        s.c1.copy_(prev.blocks[dim + 2])
        s.c2.copy_(prev.blocks[dim + 3])
        info_ = s.step()
        uv = s.u_views()
        for c in range(dim):
            u.blocks[c].copy_(uv[c])
        u.blocks[dim].copy_(s.p)
        u.blocks[dim + 1].copy_(s.c0)
        u.blocks[dim + 2].copy_(s.c1)
        u.blocks[dim + 3].copy_(s.c2)
        return info_
    raise NotImplementedError("solve: unknown form kind %r" % kind)
