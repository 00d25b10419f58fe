"""ctypes view of include/icsb200.h.

`Api(lib, prefix)` binds the C entry points of a shared library that implements the icsb200 C-ABI.
The product binds `libicsb200.so` with prefix ``icsb200_`` (icsfoam_b200/context.py); the test-only CPU
oracle exports the same signatures with prefix ``orc_`` (oracle/pyoracle.py).  Nothing in this file
computes anything.
"""
import ctypes as C

import numpy as np

# --- enums (include/icsb200.h) ---
PATCH, WALL, EMPTY, SYMMETRYPLANE, CYCLIC, PROCESSOR, CYCLICAMI = range(7)
FLUX_HLLC, FLUX_ROE, FLUX_AUSMPLUSUP = range(3)
FLUX_RUSANOV = 3   # not a reference scheme (src/Make/files:47-51): the local Lax-Friedrichs flux the brief names
FLUX_NAMES = {"HLLC": FLUX_HLLC, "ROE": FLUX_ROE, "AUSMPlusUp": FLUX_AUSMPLUSUP, "Rusanov": FLUX_RUSANOV}
LIM_UPWIND, LIM_VANLEER, LIM_MINMOD, LIM_LINEAR = range(4)
LIM_NAMES = {"upwind": LIM_UPWIND, "vanLeer": LIM_VANLEER, "Minmod": LIM_MINMOD, "linear": LIM_LINEAR}
DDT_STEADY, DDT_EULER, DDT_BACKWARD = range(3)
DDT_NAMES = {"steadyState": DDT_STEADY, "Euler": DDT_EULER, "backward": DDT_BACKWARD}
SOLVER_GMRES, SOLVER_SMOOTH = range(2)
SOLVER_NAMES = {"GMRES": SOLVER_GMRES, "smoothSolverCoupled": SOLVER_SMOOTH}
PRECOND_LUSGS, PRECOND_JACOBI = range(2)
PRECOND_NAMES = {"LUSGS": PRECOND_LUSGS, "Jacobi": PRECOND_JACOBI}
(BC_ZEROGRADIENT, BC_FIXEDVALUE, BC_SLIP, BC_EMPTY, BC_INLETOUTLET, BC_TOTALPRESSURE, BC_TOTALTEMPERATURE,
 BC_PRESSUREINLETOUTLETVELOCITY, BC_FREESTREAMPRESSURE, BC_COUPLED) = range(10)
BC_NAMES = {
    "zeroGradient": BC_ZEROGRADIENT, "fixedValue": BC_FIXEDVALUE, "slip": BC_SLIP, "symmetryPlane": BC_SLIP,
    "empty": BC_EMPTY, "inletOutlet": BC_INLETOUTLET, "freestream": BC_INLETOUTLET,
    "totalPressure": BC_TOTALPRESSURE, "totalTemperature": BC_TOTALTEMPERATURE,
    "pressureInletOutletVelocity": BC_PRESSUREINLETOUTLETVELOCITY, "freestreamPressure": BC_FREESTREAMPRESSURE,
}
FIELD_P, FIELD_U, FIELD_T = range(3)
GREAT = 1e15
SMALL = 1e-15


class Patch(C.Structure):
    _fields_ = [("kind", C.c_int), ("start", C.c_int), ("size", C.c_int), ("nbr_rank", C.c_int),
                ("nbr_patch", C.c_int), ("forwardT", C.c_double * 9)]


class Schemes(C.Structure):
    _fields_ = [("flux_scheme", C.c_int), ("limiter_rho", C.c_int), ("limiter_U", C.c_int), ("limiter_T", C.c_int),
                ("low_mach_ausm", C.c_int), ("entropy_fix_coeff", C.c_double), ("ddt_scheme", C.c_int),
                ("delta_t", C.c_double), ("local_timestepping", C.c_int), ("local_timestepping_bounding", C.c_int),
                ("local_timestepping_lower_bound", C.c_double), ("pseudo_co_num", C.c_double),
                ("pseudo_co_num_min", C.c_double), ("pseudo_co_num_max", C.c_double),
                ("pseudo_co_num_max_incr", C.c_double), ("pseudo_co_num_min_decr", C.c_double),
                ("rho_min", C.c_double), ("T_min", C.c_double), ("T_max", C.c_double), ("viscous_full_jacobian", C.c_int)]


def default_schemes(**kw):
    """Defaults as the reference reads them (initialise.H:39-71, beginTimeStep.H:8-45, updateFields.H:11-35)."""
    s = Schemes(flux_scheme=FLUX_HLLC, limiter_rho=LIM_VANLEER, limiter_U=LIM_VANLEER, limiter_T=LIM_VANLEER,
                low_mach_ausm=1, entropy_fix_coeff=0.05, ddt_scheme=DDT_STEADY, delta_t=1.0,
                local_timestepping=1, local_timestepping_bounding=1, local_timestepping_lower_bound=0.95,
                pseudo_co_num=1.0, pseudo_co_num_min=0.1, pseudo_co_num_max=25.0, pseudo_co_num_max_incr=1.25,
                pseudo_co_num_min_decr=0.1, rho_min=-GREAT, T_min=SMALL, T_max=GREAT, viscous_full_jacobian=0)
    for k, v in kw.items():
        if k == "flux_scheme" and isinstance(v, str):
            v = FLUX_NAMES[v]
        elif k.startswith("limiter") and isinstance(v, str):
            v = LIM_NAMES[v]
        elif k == "ddt_scheme" and isinstance(v, str):
            v = DDT_NAMES[v]
        setattr(s, k, v)
    return s


class SolverControls(C.Structure):
    _fields_ = [("solver", C.c_int), ("preconditioner", C.c_int), ("n_directions", C.c_int), ("max_iter", C.c_int),
                ("min_iter", C.c_int), ("tolerance", C.c_double), ("rel_tol", C.c_double)]


def solver_controls(preconditioner="LUSGS", n_directions=5, max_iter=1000, min_iter=0, tolerance=1e-12, rel_tol=1e-2, solver="GMRES",
                    n_sweeps=None):
    """fvSolution/flowSolver.  solver "smoothSolverCoupled" (smoother Jacobi): n_sweeps travels in the n_directions slot."""
    if isinstance(preconditioner, str):
        preconditioner = PRECOND_NAMES[preconditioner]
    if isinstance(solver, str):
        solver = SOLVER_NAMES[solver]
    if solver == SOLVER_SMOOTH:
        n_directions = n_sweeps if n_sweeps is not None else 1
    return SolverControls(solver, preconditioner, n_directions, max_iter, min_iter, tolerance, rel_tol)


class Residuals(C.Structure):
    _fields_ = [("s_init", C.c_double * 2), ("v_init", C.c_double * 3), ("s_final", C.c_double * 2),
                ("v_final", C.c_double * 3), ("n_iterations", C.c_int)]

    def as_dict(self):
        return {"s_init": list(self.s_init), "v_init": list(self.v_init), "s_final": list(self.s_final),
                "v_final": list(self.v_final), "n_iterations": self.n_iterations}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def dptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


# name -> (restype, argtypes) for every entry point both implementations share
SHARED_SIGNATURES = {
    "destroy": (C.c_int, [C.c_void_p]),
    "last_error": (C.c_char_p, [C.c_void_p]),
    "mesh_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                           C.c_int, C.POINTER(Patch), _ip]),
    "ami_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp]),
    "thermo_set": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]),
    "schemes_set": (C.c_int, [C.c_void_p, C.POINTER(Schemes)]),
    "bc_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "bc_set_nonuniform": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, C.c_int]),
    "state_set": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "state_get": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]),
    "boundary_get": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "new_time_step": (C.c_int, [C.c_void_p]),
    "calc_flux": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "residual": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "pseudo_dt": (C.c_int, [C.c_void_p, _dp, _dp]),
    "assemble": (C.c_int, [C.c_void_p]),
    "matrix_get_ldu": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp]),
    "matrix_set_ldu": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp]),
    "matrix_get_interfaces": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "matrix_set_interfaces": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "source_set": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "source_get": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "mrf_set": (C.c_int, [C.c_void_p, _dp, _dp]),
    "transport_set": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "matrix_mul": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp]),
    "precondition": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp]),
    "solve_delta": (C.c_int, [C.c_void_p, C.POINTER(SolverControls), _dp, _dp, _dp, C.POINTER(Residuals)]),
    "update_fields": (C.c_int, [C.c_void_p]),
    "iterate_dev": (C.c_int, [C.c_void_p, C.POINTER(SolverControls), C.POINTER(Residuals)]),
}

# entry points only the product exports
PRODUCT_SIGNATURES = {
    "create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "nccl_unique_id": (C.c_int, [C.c_void_p]),
    "iterate_host": (C.c_int, [C.c_void_p, C.POINTER(SolverControls), _dp, _dp, _dp, C.POINTER(Residuals)]),
    "launch_count": (C.c_longlong, [C.c_void_p]),
    "timers_get": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, _dp, C.POINTER(C.c_longlong), C.c_int]),
    "timers_reset": (C.c_int, [C.c_void_p, C.c_int]),
    "schedule_info": (C.c_int, [C.c_void_p, _ip]),
    "timer_begin": (C.c_int, [C.c_void_p]),
    "timer_end": (C.c_int, [C.c_void_p, _dp]),
    "hb_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _ip, _ip, _dp, _dp]),
    "phaselag_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp]),
    "hb_residuals_get": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
}

BLOCK_NC = [1, 1, 1, 1, 3, 3, 3, 3, 9]
BLOCK_NAMES = ["dSByS(0,0)", "dSByS(0,1)", "dSByS(1,0)", "dSByS(1,1)", "dSByV(0,0)", "dSByV(1,0)", "dVByS(0,0)",
               "dVByS(0,1)", "dVByV(0,0)"]


class ApiError(RuntimeError):
    pass


class Api:
    """Thin object wrapper over one context of a library implementing the icsb200 C-ABI."""

    def __init__(self, lib, prefix, signatures):
        self._lib = lib
        self._prefix = prefix
        self._fn = {}
        for name, (res, args) in signatures.items():
            f = getattr(lib, prefix + name)
            f.restype = res
            f.argtypes = args
            self._fn[name] = f
        self.h = C.c_void_p()
        self.mesh = None

    def _call(self, name, *args):
        rc = self._fn[name](self.h, *args)
        if rc != 0:
            msg = self._fn["last_error"](self.h)
            raise ApiError(f"{self._prefix}{name} failed ({rc}): {msg.decode() if msg else ''}")

    # ---- setup
    def mesh_set(self, mesh):
        self.mesh = mesh
        patches = (Patch * len(mesh.patches))()
        for i, p in enumerate(mesh.patches):
            patches[i].kind, patches[i].start, patches[i].size = p["kind"], p["start"], p["size"]
            patches[i].nbr_rank, patches[i].nbr_patch = p.get("nbr_rank", -1), p.get("nbr_patch", -1)
            for k, v in enumerate(p.get("forwardT", [1, 0, 0, 0, 1, 0, 0, 0, 1])):
                patches[i].forwardT[k] = v
        for i, p in enumerate(mesh.patches):
            if p["kind"] == CYCLICAMI and "ami" in p:   # AMI addressing / weights must be known when the mesh is set
                start, face, weight = p["ami"]
                self._call("ami_set", i, p["size"], iptr(np.ascontiguousarray(start, np.int32)), iptr(np.ascontiguousarray(face, np.int32)),
                           dptr(np.ascontiguousarray(weight, np.float64)))
        sd = np.asarray(mesh.solutionD, dtype=np.int32)
        self._call("mesh_set", mesh.n_cells, mesh.n_internal_faces, mesh.n_faces, iptr(mesh.owner), iptr(mesh.neighbour),
                   dptr(mesh.Sf), dptr(mesh.magSf), dptr(mesh.weights), dptr(mesh.deltaCoeffs),
                   dptr(mesh.nonOrthDeltaCoeffs), dptr(mesh.C), dptr(mesh.V), dptr(mesh.Cf), len(mesh.patches), patches,
                   iptr(sd))

    def thermo_set(self, R, Cp, mu=0.0, Pr=1.0):
        self._call("thermo_set", R, Cp, mu, Pr)

    def schemes_set(self, s):
        self.schemes = s
        self._call("schemes_set", C.byref(s))

    def bc_set(self, patch, field, kind, params=()):
        if isinstance(patch, str):
            patch = [p["name"] for p in self.mesh.patches].index(patch)
        if isinstance(kind, str):
            kind = BC_NAMES[kind]
        if isinstance(params, np.ndarray) and params.ndim == 2:      # one row per face: nonuniform List<...>
            prm = np.ascontiguousarray(params, np.float64)
            assert prm.shape[0] == self.mesh.patches[patch]["size"]
            self._call("bc_set_nonuniform", patch, field, kind, dptr(prm), int(prm.shape[1]))
            return
        prm = np.asarray(list(params), dtype=np.float64)
        self._call("bc_set", patch, field, kind, dptr(prm) if prm.size else None, int(prm.size))

    # ---- state
    def state_set(self, p, U, T):
        self._call("state_set", dptr(np.ascontiguousarray(p, np.float64)), dptr(np.ascontiguousarray(U, np.float64)),
                   dptr(np.ascontiguousarray(T, np.float64)))

    def state_get(self):
        N = self.mesh.n_cells
        out = {"rho": np.empty(N), "rhoU": np.empty((N, 3)), "rhoE": np.empty(N), "p": np.empty(N), "U": np.empty((N, 3)),
               "T": np.empty(N)}
        self._call("state_get", *[dptr(out[k]) for k in ("rho", "rhoU", "rhoE", "p", "U", "T")])
        return out

    def boundary_get(self):
        NB = self.mesh.n_faces - self.mesh.n_internal_faces
        out = {"rho": np.empty(NB), "U": np.empty((NB, 3)), "p": np.empty(NB), "T": np.empty(NB)}
        self._call("boundary_get", *[dptr(out[k]) for k in ("rho", "U", "p", "T")])
        return out

    def new_time_step(self):
        self._call("new_time_step")

    # ---- hot path, piecewise
    def calc_flux(self):
        FT = self.mesh.n_faces
        phi, phiUp, phiEp = np.zeros(FT), np.zeros((FT, 3)), np.zeros(FT)
        self._call("calc_flux", dptr(phi), dptr(phiUp), dptr(phiEp))
        return phi, phiUp, phiEp

    def residual(self):
        N = self.mesh.n_cells
        a, b, c = np.zeros(N), np.zeros((N, 3)), np.zeros(N)
        self._call("residual", dptr(a), dptr(b), dptr(c))
        return a, b, c

    def pseudo_dt(self):
        N = self.mesh.n_cells
        a, b = np.zeros(N), np.zeros(N)
        self._call("pseudo_dt", dptr(a), dptr(b))
        return a, b

    def assemble(self):
        self._call("assemble")

    def matrix_get_ldu(self, block):
        nc, N, F = BLOCK_NC[block], self.mesh.n_cells, self.mesh.n_internal_faces
        d, u, l = np.zeros((N, nc)), np.zeros((F, nc)), np.zeros((F, nc))
        self._call("matrix_get_ldu", block, dptr(d), dptr(u), dptr(l))
        return d, u, l

    def matrix_set_ldu(self, block, diag, upper=None, lower=None):
        self._call("matrix_set_ldu", block, dptr(np.ascontiguousarray(diag)),
                   dptr(np.ascontiguousarray(upper)) if upper is not None else None,
                   dptr(np.ascontiguousarray(lower)) if lower is not None else None)

    def matrix_get_interfaces(self, block):
        NB = self.mesh.n_faces - self.mesh.n_internal_faces
        a = np.zeros((NB, BLOCK_NC[block]))
        self._call("matrix_get_interfaces", block, dptr(a))
        return a

    def matrix_set_interfaces(self, block, int_upper):
        self._call("matrix_set_interfaces", block, dptr(np.ascontiguousarray(int_upper, np.float64)))

    def source_set(self, sRho, sRhoU, sRhoE):
        self._call("source_set", dptr(np.ascontiguousarray(sRho)), dptr(np.ascontiguousarray(sRhoU)),
                   dptr(np.ascontiguousarray(sRhoE)))

    def source_get(self):
        """sources of the assembled system as the solver sees them (R*V, HB and MRF terms)"""
        N = self.mesh.n_cells
        a, b, c = np.zeros(N), np.zeros((N, 3)), np.zeros(N)
        self._call("source_get", dptr(a), dptr(b), dptr(c))
        return a, b, c

    def mrf_set(self, face_velocity=None, omega=None):
        """flux.MRFFaceVelocity() [n_faces] and flux.MRFOmega() [n_cells,3] (outerLoop.H:18-21); None = zero field"""
        fv = None if face_velocity is None else np.ascontiguousarray(face_velocity, np.float64)
        om = None if omega is None else np.ascontiguousarray(omega, np.float64)
        assert fv is None or fv.size == self.mesh.n_faces
        assert om is None or om.size == 3 * self.mesh.n_cells
        self._call("mrf_set", dptr(fv), dptr(om))

    def transport_set(self, muEff=None, muEff_b=None, alphaEff=None, alphaEff_b=None):
        """turbulence->muEff() / alphaEff(): cell [n_cells] and boundary-face [n_faces - n_internal_faces] values; None = laminar"""
        if muEff is None:
            self._call("transport_set", None, None, None, None)
            return
        arrs = [np.ascontiguousarray(a, np.float64) for a in (muEff, muEff_b, alphaEff, alphaEff_b)]
        NB = self.mesh.n_faces - self.mesh.n_internal_faces
        assert arrs[0].size == arrs[2].size == self.mesh.n_cells and arrs[1].size == arrs[3].size == NB
        self._call("transport_set", *[dptr(a) for a in arrs])

    def matrix_mul(self, xRho, xRhoU, xRhoE):
        N = self.mesh.n_cells
        a, b, c = np.zeros(N), np.zeros((N, 3)), np.zeros(N)
        self._call("matrix_mul", dptr(np.ascontiguousarray(xRho)), dptr(np.ascontiguousarray(xRhoU)),
                   dptr(np.ascontiguousarray(xRhoE)), dptr(a), dptr(b), dptr(c))
        return a, b, c

    def precondition(self, kind, xRho, xRhoU, xRhoE):
        if isinstance(kind, str):
            kind = PRECOND_NAMES[kind]
        a, b, c = (np.array(xRho, dtype=np.float64, order="C"), np.array(xRhoU, dtype=np.float64, order="C"),
                   np.array(xRhoE, dtype=np.float64, order="C"))
        self._call("precondition", kind, dptr(a), dptr(b), dptr(c))
        return a, b, c

    def solve_delta(self, ctl):
        N = self.mesh.n_cells
        a, b, c = np.zeros(N), np.zeros((N, 3)), np.zeros(N)
        res = Residuals()
        self._call("solve_delta", C.byref(ctl), dptr(a), dptr(b), dptr(c), C.byref(res))
        return (a, b, c), res

    def update_fields(self):
        self._call("update_fields")

    def iterate(self, ctl):
        res = Residuals()
        self._call("iterate_dev", C.byref(ctl), C.byref(res))
        return res

    def close(self):
        if self.h:
            self._fn["destroy"](self.h)
            self.h = C.c_void_p()
