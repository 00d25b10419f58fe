"""The named configurations of BASELINE.json as concrete inputs.

A `Case` bundles a mesh, thermo, schemes, solver controls, boundary conditions and initial fields, and can
`apply()` itself to any object exposing the icsb200 C-ABI wrapper (`capi.Api`): the CUDA product
(`icsfoam_b200.context.Context`) or the test-only oracle.  Dictionary key names follow the tutorial
dictionaries (fvSchemes / fvSolution / thermophysicalProperties / 0/*).
"""
import os

import numpy as np

from . import capi
from . import meshtools as mt

RR = 8314.46261815324  # J/(kmol K): OpenFOAM thermodynamic::RR = 1e3 * physicoChemical::R (kept a parameter)


class Case:
    def __init__(self, name, mesh, R, Cp, schemes, controls, bcs, p, U, T, mu=0.0, Pr=1.0, n_iter_default=20):
        self.name, self.mesh, self.R, self.Cp, self.mu, self.Pr = name, mesh, R, Cp, mu, Pr
        self.schemes, self.controls, self.bcs = schemes, controls, bcs
        self.p = np.ascontiguousarray(p, np.float64)
        self.U = np.ascontiguousarray(U, np.float64)
        self.T = np.ascontiguousarray(T, np.float64)
        self.n_iter_default = n_iter_default
        self.mrf = None   # (omega[3], origin[3], translation velocity[3], zone predicate on points or None): see with_mrf
        self.transport = None  # callable(points[n,3]) -> (muEff, alphaEff): stands in for a turbulence model's fields

    def with_mrf(self, omega=(0, 0, 0), origin=(0, 0, 0), velocity=(0, 0, 0), zone=None):
        """Frame motion of an MRFCoupledZone (rotation `omega` about `origin`) plus an MRFTranslatingZone (`velocity`),
        restricted to the points where zone(xyz) is true (None: whole mesh)."""
        self.mrf = (np.asarray(omega, float), np.asarray(origin, float), np.asarray(velocity, float), zone)
        return self

    def with_transport(self, fn):
        """Effective viscosity / thermal diffusivity fields as a turbulence model would hand them over every outer iteration
        (turbulence->muEff(), alphaEff()): fn(points) -> (muEff, alphaEff), evaluated at cell and boundary-face centres."""
        self.transport = fn
        return self

    def transport_fields(self, m):
        F = m.n_internal_faces
        mu_c, al_c = self.transport(m.C)
        mu_b, al_b = self.transport(m.Cf[F:])
        return tuple(np.ascontiguousarray(a, np.float64) for a in (mu_c, mu_b, al_c, al_b))

    def mrf_fields(self, m):
        """flux.MRFFaceVelocity() = (MRF.faceU() + MRFTrans.faceU()) & Sf/magSf and flux.MRFOmega() of mesh `m`
        (outerLoop.H:18-21; host set-up, src/cfdTools/MRFCoupled/MRFCoupledZone.C faceU / omega)."""
        omega, origin, vel, zone = self.mrf
        faceU = np.cross(omega, m.Cf - origin) + vel
        om = np.tile(omega, (m.n_cells, 1))
        if zone is not None:
            faceU[~zone(m.Cf)] = 0.0
            om[~zone(m.C)] = 0.0
        fv = (faceU * (m.Sf / m.magSf[:, None])).sum(1)
        for p in m.patches:
            if p["kind"] == capi.EMPTY:
                fv[p["start"]:p["start"] + p["size"]] = 0.0
        return np.ascontiguousarray(fv), np.ascontiguousarray(om)

    def apply(self, api, mesh=None, cells=None):
        """Configure `api` with this case.  `mesh`/`cells` select a partition (cells = global cell ids)."""
        m = mesh or self.mesh
        api.mesh_set(m)
        api.thermo_set(self.R, self.Cp, self.mu, self.Pr)
        api.schemes_set(self.schemes)
        names = [p["name"] for p in m.patches]
        for patch, fields in self.bcs.items():
            if patch not in names:
                continue
            for field, (kind, params) in fields.items():
                if isinstance(params, np.ndarray) and params.ndim == 2 and m is not self.mesh:
                    # non-uniform entries on a partition: the rows of the faces this rank keeps
                    rows = m.face_global[m.patch_faces(patch)] - self.mesh.patches[self.mesh.patch_index(patch)]["start"]
                    params = np.ascontiguousarray(params[rows])
                api.bc_set(patch, {"p": capi.FIELD_P, "U": capi.FIELD_U, "T": capi.FIELD_T}[field], kind, params)
        if self.mrf is not None:
            api.mrf_set(*self.mrf_fields(m))
        if self.transport is not None:
            api.transport_set(*self.transport_fields(m))
        if cells is None:
            api.state_set(self.p, self.U, self.T)
        else:
            api.state_set(self.p[cells], self.U[cells], self.T[cells])
        return api

    def partition(self, n_parts, mode="x", only=None):
        """Slab / block decomposition of the structured mesh (decomposePar 'simple' stand-in).  `only`: extract just that
        rank's sub-mesh (the other list entries are None)."""
        C = self.mesh.C
        if mode == "x":
            order = np.argsort(C[:, 0], kind="stable")
            part = np.empty(self.mesh.n_cells, np.int32)
            part[order] = (np.arange(self.mesh.n_cells) * n_parts // self.mesh.n_cells).astype(np.int32)
        else:  # 2x2x2-style blocks by coordinate medians
            nx, ny, nz = mode
            part = np.zeros(self.mesh.n_cells, np.int32)
            stride = 1
            for d, n in enumerate((nx, ny, nz)):       # rank = ix + nx*(iy + ny*iz), as meshtools.structured_part
                q = np.quantile(C[:, d], np.linspace(0, 1, n + 1)[1:-1]) if n > 1 else []
                part += stride * np.searchsorted(q, C[:, d]).astype(np.int32)
                stride *= n
        # decomposeParDict `preservePatches` (VKI-LS89/system/decomposeParDict:27-35): both cells of every cyclic face pair on one rank
        for i, pa in enumerate(self.mesh.patches):
            if pa["kind"] == capi.CYCLIC and i < pa["nbr_patch"]:
                pb = self.mesh.patches[pa["nbr_patch"]]
                fa = np.arange(pa["start"], pa["start"] + pa["size"])
                fb = np.arange(pb["start"], pb["start"] + pb["size"])
                part[self.mesh.owner[fb]] = part[self.mesh.owner[fa]]
        meshes = [self.mesh.extract_part(part, r) if only is None or r == only else None for r in range(n_parts)]
        return part, meshes

    def decomposed(self, case_dir):
        """The partition OpenFOAM's decomposePar wrote under `case_dir/processor*`: same return value as
        `partition`.  `self.mesh` must be the undecomposed mesh of that case (cellProcAddressing refers to it)."""
        from .meshtools import foamcase
        n = foamcase.n_processors(case_dir)
        if n == 0:
            raise RuntimeError(f"no processor directories under {case_dir}")
        cache = {}
        meshes = [foamcase.read_decomposed(case_dir, r, cache) for r in range(n)]
        part = np.full(self.mesh.n_cells, -1, np.int32)
        for r, m in enumerate(meshes):
            if m.cell_global is None:
                raise RuntimeError(f"processor{r} has no cellProcAddressing")
            part[m.cell_global] = r
        if (part < 0).any():
            raise RuntimeError("the processor directories do not cover the mesh")
        return part, meshes



_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_tutorials():
    """Directory holding the reference's tutorials (input data): $ICSFOAM_REF/tutorials, else /root/reference/tutorials."""
    for base in (os.environ.get("ICSFOAM_REF"), "/root/reference"):
        if base and os.path.isdir(os.path.join(base, "tutorials")):
            return os.path.join(base, "tutorials")
    return None


def tutorial_dir(name):
    """Case directory of a shipped tutorial (`forwardStep`, `VKI-LS89`): the reference checkout when this machine has one,
    else the copy `stage_tutorials()` left under cases_local/ (a GPU box has no /root/reference).  None if neither exists."""
    ref = reference_tutorials()
    for d in ((os.path.join(ref, name) if ref else None), os.path.join(_ROOT, "cases_local", name)):
        if d and os.path.isdir(os.path.join(d, "constant", "polyMesh")):
            return d
    return None


def stage_tutorials(names=("forwardStep", "VKI-LS89")):
    """Copy the tutorial directories the C2 / C5 tests run on from the reference checkout into cases_local/ (git-ignored
    input data — meshes and dictionaries, no sources — that travels to the GPU box with the working tree).  Called by
    __graft_entry__.build(); a no-op on a machine without the reference."""
    import shutil
    ref = reference_tutorials()
    if not ref:
        return []
    done = []
    for n in names:
        src, dst = os.path.join(ref, n), os.path.join(_ROOT, "cases_local", n)
        if os.path.isdir(src) and not os.path.isdir(os.path.join(dst, "constant", "polyMesh")):
            shutil.copytree(src, dst, dirs_exist_ok=True)
            for r, ds, fs in os.walk(dst):   # the reference tree is read-only; the copy must be removable
                for x in ds + fs:
                    os.chmod(os.path.join(r, x), 0o755 if x in ds or os.access(os.path.join(r, x), os.X_OK) else 0o644)
            done.append(n)
    return done

def _uniform(mesh, p, U, T):
    N = mesh.n_cells
    return np.full(N, float(p)), np.tile(np.asarray(U, float), (N, 1)), np.full(N, float(T))


def shock_tube(n=500, flux="ROE", transient=True):
    """C1 tutorials/shockTube: 1-D Sod tube.  ROE as shipped (fvSchemes:19-33) or AUSMPlusUp (BASELINE.json)."""
    mesh = mt.shock_tube(n)
    R, Cp = RR / 28.96, 1004.5
    p, U, T = _uniform(mesh, 1e5, (0, 0, 0), 348.432)
    right = (mesh.C[:, 0] >= 0.0)  # setFieldsDict:19-37 boxToCell (0 -1 -1) (0.5 1 1)
    p[right], T[right] = 1e4, 278.746
    sch = capi.default_schemes(flux_scheme=flux, limiter_rho="vanLeer", limiter_U="vanLeer", limiter_T="vanLeer",
                               ddt_scheme="backward" if transient else "steadyState", delta_t=1e-6, pseudo_co_num=1.0)
    ctl = capi.solver_controls("LUSGS", n_directions=5, max_iter=20, tolerance=1e-12, rel_tol=1e-4)
    zg = ("zeroGradient", ())
    bcs = {n_: {"p": zg, "T": zg, "U": zg} for n_ in ("side1", "side2")}
    for n_ in ("wallYmin", "wallYmax", "wallZmin", "wallZmax"):
        bcs[n_] = {"p": zg, "T": zg, "U": ("slip", ())}
    return Case("shockTube", mesh, R, Cp, sch, ctl, bcs, p, U, T, n_iter_default=10)


def bump(nxb=66, ny=54, co=200.0, mu=0.0, Pr=0.71):
    """C3 tutorials/circularArcBump/transonic, optionally refined (bump-4M: nxb=1280, ny=1040)."""
    mesh = mt.bump(nxb, ny)
    R, Cp = RR / 28.966, 1005.0
    p, U, T = _uniform(mesh, 74653.0, (218.0, 0, 0), 274.9)
    sch = capi.default_schemes(flux_scheme="HLLC", limiter_rho="Minmod", limiter_U="Minmod", limiter_T="Minmod",
                               ddt_scheme="steadyState", pseudo_co_num=co, pseudo_co_num_max=co)
    ctl = capi.solver_controls("LUSGS", n_directions=5, max_iter=10, tolerance=1e-10, rel_tol=1e-2)
    zg = ("zeroGradient", ())
    bcs = {
        "INLE1": {"p": ("totalPressure", (101300.0, 1.4)), "U": ("pressureInletOutletVelocity", (0, 0, 0)),
                  "T": ("totalTemperature", (288.15, 1.4))},
        "PRES2": {"p": ("fixedValue", (74653.0,)), "U": zg, "T": zg},
        "WALL3": {"p": zg, "U": ("slip", ()), "T": zg},
        "WALL4": {"p": zg, "U": ("slip", ()), "T": zg},
    }
    return Case("bump", mesh, R, Cp, sch, ctl, bcs, p, U, T, mu=mu, Pr=Pr, n_iter_default=20)


def onera_box(n=48, co=100.0, flux="HLLC", parts=None, rank=0, mu=0.0, Pr=0.71):
    """C4 synthetic stand-in for tutorials/OneraM6Wing (inviscid, HLLC, vanLeer, steady, Co=100).
    With parts=(px,py,pz) only partition `rank` is generated (fields are uniform, so no global arrays are needed)."""
    if parts is None:
        mesh = mt.onera_box(n)
    else:
        mesh = mt.structured_part((n, n, n), parts, rank, 2, (-1.0, 0.0, 0.0), (2.0, 3.0, 3.0), amp=0.12,
                                  patch_kinds=(mt.PATCH, mt.PATCH, mt.SYMMETRYPLANE, mt.PATCH, mt.WALL, mt.PATCH),
                                  patch_names=("inlet", "outlet", "symmetry", "lateral", "wing", "top"))
    R, Cp = RR / 28.966, 1005.0
    Uinf = (285.6, 15.268, 0.0)
    p, U, T = _uniform(mesh, 101325.0, Uinf, 288.15)
    sch = capi.default_schemes(flux_scheme=flux, limiter_rho="vanLeer", limiter_U="vanLeer", limiter_T="vanLeer",
                               ddt_scheme="steadyState", pseudo_co_num=co, pseudo_co_num_max=co)
    ctl = capi.solver_controls("LUSGS", n_directions=5, max_iter=10, tolerance=1e-6, rel_tol=1e-1)
    slip = ("slip", ())
    fs = {"p": ("freestreamPressure", (101325.0,) + Uinf), "U": ("freestream", Uinf), "T": ("inletOutlet", (288.15,))}
    bcs = {"wing": {"p": slip, "U": slip, "T": slip}, "symmetry": {"p": slip, "U": slip, "T": slip}}
    for n_ in ("inlet", "outlet", "lateral", "top"):
        bcs[n_] = dict(fs)
    return Case("onera-box", mesh, R, Cp, sch, ctl, bcs, p, U, T, mu=mu, Pr=Pr, n_iter_default=20)


def periodic_box(n=8, flux="HLLC", limiter="vanLeer", seed=0, nz=None, cyclic=True, mu=0.0, Pr=0.71, ami_shift=None):
    """Small randomised box with a translational cyclic pair in x — a parity-test workhorse, not a tutorial."""
    mesh = mt.structured(1, n, n, nz or n, 0, (0, 0, 0), (1.0, 1.2, 0.9),
                         patch_kinds=(capi.PATCH, capi.PATCH, capi.WALL, capi.PATCH, capi.SYMMETRYPLANE, capi.PATCH))
    if cyclic and ami_shift is not None:
        mesh.set_cyclic_ami("xmin", "xmax", ami_shift)   # non-conformal periodic pair (cyclicAMI)
    elif cyclic:
        mesh.set_cyclic("xmin", "xmax")
    rng = np.random.default_rng(seed)
    N = mesh.n_cells
    p = 1e5 * (1 + 0.2 * rng.random(N))
    T = 300 * (1 + 0.2 * rng.random(N))
    U = 150 * (rng.random((N, 3)) - 0.3)
    sch = capi.default_schemes(flux_scheme=flux, limiter_rho=limiter, limiter_U=limiter, limiter_T=limiter,
                               ddt_scheme="steadyState", pseudo_co_num=5.0, pseudo_co_num_max=50.0)
    ctl = capi.solver_controls("LUSGS", n_directions=5, max_iter=10, tolerance=1e-10, rel_tol=1e-3)
    zg, slip = ("zeroGradient", ()), ("slip", ())
    bcs = {
        "ymin": {"p": zg, "U": slip, "T": zg},
        "ymax": {"p": ("fixedValue", (1.05e5,)), "U": ("inletOutlet", (50.0, 10.0, 0.0)), "T": ("inletOutlet", (310.0,))},
        "zmin": {"p": slip, "U": slip, "T": slip},
        "zmax": {"p": ("freestreamPressure", (1.0e5, 50.0, 10.0, 20.0)), "U": ("freestream", (50.0, 10.0, 20.0)),
                 "T": ("fixedValue", (305.0,))},
    }
    return Case("periodic-box", mesh, 287.0, 1005.0, sch, ctl, bcs, p, U, T, mu=mu, Pr=Pr)


def rot_box(n=6, flux="HLLC", limiter="vanLeer", seed=0, nz=None, mu=0.0, ami_shift=None):
    """A 90-degree sector: the cube [0,1]^2 x [0,0.9] whose faces x=0 and y=0 form a ROTATIONAL cyclic pair about the z axis
    through the corner (four copies tile the plane around it).  forwardT of `xmin` is the rotation by +90 degrees about z
    (it turns the outward normal -y of `ymin` into +x).  Random state: a parity workhorse for
    cyclicFvPatchField::patchNeighbourField with doTransform() — SURVEY 8f-4."""
    mesh = mt.structured(1, n, n, nz or n, 0, (0, 0, 0), (1.0, 1.0, 0.9),
                         patch_kinds=(capi.PATCH, capi.PATCH, capi.PATCH, capi.PATCH, capi.SYMMETRYPLANE, capi.PATCH))
    if ami_shift is None:
        mesh.set_cyclic_rotational("xmin", "ymin", [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    else:   # non-conformal rotational pair (rotational cyclicAMI): every face sees two rotated neighbour faces
        mesh.set_cyclic_ami_rotational("xmin", "ymin", [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], ami_shift)
    rng = np.random.default_rng(seed)
    N = mesh.n_cells
    p = 1e5 * (1 + 0.2 * rng.random(N))
    T = 300 * (1 + 0.2 * rng.random(N))
    U = 150 * (rng.random((N, 3)) - 0.3)
    sch = capi.default_schemes(flux_scheme=flux, limiter_rho=limiter, limiter_U=limiter, limiter_T=limiter,
                               ddt_scheme="steadyState", pseudo_co_num=5.0, pseudo_co_num_max=50.0)
    ctl = capi.solver_controls("LUSGS", n_directions=5, max_iter=10, tolerance=1e-10, rel_tol=1e-3)
    zg, slip = ("zeroGradient", ()), ("slip", ())
    bcs = {
        "xmax": {"p": zg, "U": slip, "T": zg},
        "ymax": {"p": ("fixedValue", (1.05e5,)), "U": ("inletOutlet", (50.0, 10.0, 0.0)), "T": ("inletOutlet", (310.0,))},
        "zmin": {"p": slip, "U": slip, "T": slip},
        "zmax": {"p": ("freestreamPressure", (1.0e5, 50.0, 10.0, 20.0)), "U": ("freestream", (50.0, 10.0, 20.0)),
                 "T": ("fixedValue", (305.0,))},
    }
    return Case("rot-box", mesh, 287.0, 1005.0, sch, ctl, bcs, p, U, T, mu=mu, Pr=0.71)


def sector_and_annulus(n=6, mu=0.0, flux="ROE", limiter="upwind", seed=3):
    """A 90-degree sector with a ROTATIONAL cyclic pair and the full annulus it stands for: the square [-1,1]^2 tiled by four
    rotated copies of the sector [0,1]^2, with the rotation-symmetric copy of the sector's (random) state and rotation-
    covariant boundary conditions (slip walls).  Returns (sector case, annulus case, annulus cells of the first quadrant,
    matching sector cells).  2-D (empty z patches).  Both runs must agree wherever the reference's treatment of the pair is
    exact: first-order reconstruction (see DESIGN Q13 for the limited one) and all viscous terms."""
    T = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    kinds = (capi.PATCH, capi.PATCH, capi.PATCH, capi.PATCH, capi.EMPTY, capi.EMPTY)
    sec = mt.structured(1, n, n, 1, 0, (0, 0, 0), (1.0, 1.0, 0.1), patch_kinds=kinds)
    sec.set_cyclic_rotational("xmin", "ymin", T)
    full = mt.structured(1, 2 * n, 2 * n, 1, 0, (-1.0, -1.0, 0), (1.0, 1.0, 0.1), patch_kinds=kinds)
    rng = np.random.default_rng(seed)
    N = sec.n_cells
    p0, T0 = 1e5 * (1 + 0.1 * rng.random(N)), 300 * (1 + 0.1 * rng.random(N))
    U0 = 80 * (rng.random((N, 3)) - 0.4)
    U0[:, 2] = 0
    Xk, turns = full.C.copy(), np.zeros(full.n_cells, int)
    for _ in range(3):                                   # rotate back by -90 degrees until the centre lies in the sector
        todo = ~((Xk[:, 0] > 0) & (Xk[:, 1] > 0))
        Xk[todo] = Xk[todo] @ T
        turns[todo] += 1
    key = lambda P: np.round(P[:, 0] * n - 0.5).astype(int) + n * np.round(P[:, 1] * n - 0.5).astype(int)
    lut = np.empty(N, int)
    lut[key(sec.C)] = np.arange(N)
    src = lut[key(Xk)]
    UF = U0[src].copy()
    for q in range(1, 4):
        UF[turns == q] = UF[turns == q] @ np.linalg.matrix_power(T, q).T
    sch = capi.default_schemes(flux_scheme=flux, limiter_rho=limiter, limiter_U=limiter, limiter_T=limiter, ddt_scheme="steadyState",
                               pseudo_co_num=5.0, pseudo_co_num_max=50.0)
    ctl = capi.solver_controls("Jacobi", n_directions=6, max_iter=40, tolerance=1e-14, rel_tol=1e-10)
    wall = {"p": ("zeroGradient", ()), "U": ("slip", ()), "T": ("zeroGradient", ())}
    cs = Case("sector", sec, 287.0, 1005.0, sch, ctl, {nm: dict(wall) for nm in ("xmax", "ymax")}, p0, U0, T0, mu=mu, Pr=0.71)
    cf = Case("annulus", full, 287.0, 1005.0, sch, ctl, {nm: dict(wall) for nm in ("xmin", "xmax", "ymin", "ymax")}, p0[src], UF, T0[src], mu=mu, Pr=0.71)
    first = np.nonzero(turns == 0)[0]
    return cs, cf, first, src[first]


def with_inlet_profiles(case, patch, p_kind="fixedValue"):
    """Replace the uniform entries of `patch` by non-uniform ones (`nonuniform List<...>`): a pressure (or total pressure)
    profile, a velocity profile for the inletOutlet `inletValue` and a temperature profile — what a radially varying
    turbomachinery inlet looks like in 0/p, 0/U, 0/T."""
    m = case.mesh
    x = m.Cf[m.patch_faces(patch)]
    s = 0.5 + 0.5 * np.sin(3.0 * x[:, 0] + 2.0 * x[:, 2])
    if p_kind == "totalPressure":
        p_rows = np.column_stack([101300.0 * (1.0 + 0.04 * s), np.full(len(s), 1.4)])
    else:
        p_rows = (1.05e5 * (1.0 + 0.03 * s))[:, None]
    case.bcs[patch] = dict(case.bcs[patch])
    case.bcs[patch]["p"] = (p_kind, np.ascontiguousarray(p_rows))
    ukind = case.bcs[patch]["U"][0]
    if ukind in ("inletOutlet", "fixedValue", "freestream"):
        case.bcs[patch]["U"] = (ukind, np.ascontiguousarray(np.column_stack([50.0 + 20.0 * s, 10.0 - 5.0 * s, 3.0 * s])))
    tkind = case.bcs[patch]["T"][0]
    if tkind == "totalTemperature":
        case.bcs[patch]["T"] = (tkind, np.ascontiguousarray(np.column_stack([288.15 * (1.0 + 0.02 * s), np.full(len(s), 1.4)])))
    else:
        case.bcs[patch]["T"] = (tkind, np.ascontiguousarray((310.0 + 8.0 * s)[:, None]))
    return case


def scrambled_box(n=6, flux="HLLC", limiter="vanLeer", seed=0, mu=0.0):
    """A box whose cells are renumbered at random: irregular LU-SGS levels, rows with up to 6 lower (or upper)
    neighbours, no structure for the tile heuristics to find — the 'unstructured numbering' stress case."""
    base = periodic_box(n, flux, limiter, seed, mu=mu)
    rng = np.random.default_rng(seed + 1000)
    perm = rng.permutation(base.mesh.n_cells).astype(np.int32)
    mesh = base.mesh.renumber(perm)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size, dtype=np.int32)
    # the cyclic pair survives renumbering only if both patches keep matching face order: re-pair by face centre
    for p in mesh.patches:
        if p["kind"] == capi.CYCLIC:
            p["kind"], p["nbr_patch"] = capi.PATCH, -1
    c = Case("scrambled-box", mesh, base.R, base.Cp, base.schemes, base.controls,
             {k: v for k, v in base.bcs.items()}, base.p[inv], base.U[inv], base.T[inv], mu=base.mu, Pr=base.Pr)
    return c


def forward_step(polymesh_dir):
    """C2 tutorials/forwardStep: Mach-3 step on the shipped polyhedral 2-D mesh (36 576 cells), HLLC, Minmod, transient
    dual time (backward, deltaT 1e-3), GMRES m=8 maxIter 20 relTol 1e-4, LU-SGS.  The mesh is read from `polymesh_dir`
    (the reference's constant/polyMesh — input data, never copied into this repository)."""
    mesh = mt.read_polymesh(polymesh_dir)
    R, Cp = RR / 11640.3, 2.5          # "normalised" gas: c = 1 m/s at T = 1 K, gamma = 7/5
    p, U, T = _uniform(mesh, 1.0, (3.0, 0, 0), 1.0)
    sch = capi.default_schemes(flux_scheme="HLLC", limiter_rho="Minmod", limiter_U="Minmod", limiter_T="Minmod",
                               ddt_scheme="backward", delta_t=1e-3, pseudo_co_num=1.0, pseudo_co_num_max=10.0)
    ctl = capi.solver_controls("LUSGS", n_directions=8, max_iter=20, tolerance=1e-12, rel_tol=1e-4)
    zg, slip = ("zeroGradient", ()), ("slip", ())
    bcs = {
        "inlet": {"p": ("fixedValue", (1.0,)), "U": ("fixedValue", (3.0, 0, 0)), "T": ("fixedValue", (1.0,))},
        "outlet": {"p": zg, "U": ("inletOutlet", (3.0, 0, 0)), "T": ("inletOutlet", (1.0,))},
        "bottom": {"p": slip, "U": slip, "T": slip},
        "top": {"p": slip, "U": slip, "T": slip},
        "obstacle": {"p": zg, "U": slip, "T": zg},
    }
    return Case("forwardStep", mesh, R, Cp, sch, ctl, bcs, p, U, T, n_iter_default=50)


def vki_ls89(polymesh_dir, mu=1.5e-5, co=10.0):
    """C5 tutorials/VKI-LS89: turbine cascade on the shipped 2-D mesh (28 059 cells, translational cyclic pair
    Upper/Lower_periodicity, no-slip isothermal blade), ROE + vanLeer, steady dual time with local time stepping
    (pseudoCoNum 10, max 50), GMRES m=8 maxIter 10 relTol 1e-3, LU-SGS (system/fvSchemes, system/fvSolution, 0/p 0/U 0/T,
    constant/thermophysicalProperties).  The tutorial runs RAS kOmegaSST; here muEff = mu (laminar).
    The mesh is read from `polymesh_dir` (the reference's constant/polyMesh — input data, never copied into this repository)."""
    mesh = mt.read_polymesh(polymesh_dir)
    R, Cp = RR / 28.966, 1005.0
    p, U, T = _uniform(mesh, 1e5, (100.0, 0, 0), 400.0)
    sch = capi.default_schemes(flux_scheme="ROE", limiter_rho="vanLeer", limiter_U="vanLeer", limiter_T="vanLeer", entropy_fix_coeff=0.05,
                               ddt_scheme="steadyState", pseudo_co_num=co, pseudo_co_num_max=50.0)
    ctl = capi.solver_controls("LUSGS", n_directions=8, max_iter=10, tolerance=1e-12, rel_tol=1e-3)
    zg = ("zeroGradient", ())
    piov = ("pressureInletOutletVelocity", (0, 0, 0))
    bcs = {
        "inlet": {"p": ("totalPressure", (160500.0, 1.4)), "U": piov, "T": ("totalTemperature", (420.0, 1.4))},
        "outlet": {"p": ("fixedValue", (82000.0,)), "U": piov, "T": ("inletOutlet", (400.0,))},
        "blade": {"p": zg, "U": ("fixedValue", (0, 0, 0)), "T": ("fixedValue", (301.0,))},     # noSlip
    }
    return Case("VKI-LS89", mesh, R, Cp, sch, ctl, bcs, p, U, T, mu=mu, Pr=0.72, n_iter_default=20)


class HBCase:
    """Harmonic Balance configuration (C5 'vki-HB' stand-in; dbnsFullyImplicitHBFoam): one base `Case` per time instance
    (same mesh, instance-specific boundary values and initial fields) plus the HBZone set-up.

    `apply(api)` configures the product through the icsb200 C-ABI on the instance-replicated mesh
    (`hb.ReplicatedMesh` + `icsb200_hb_set`); `instances` are what a per-instance implementation (the oracle) consumes."""

    def __init__(self, name, instances, snapshots, D, zone_of_cell=None, cyl_coords=None, rotation_axis=None, rotation_centre=None):
        from . import hb
        self.name, self.instances, self.snapshots = name, instances, np.asarray(snapshots, float)
        self.n_instants = len(instances)
        self.D = np.ascontiguousarray(np.asarray(D, float).reshape(-1, self.n_instants, self.n_instants))
        self.zone_of_cell = None if zone_of_cell is None else np.ascontiguousarray(zone_of_cell, np.int32)
        self.cyl_coords, self.rotation_axis, self.rotation_centre = cyl_coords, rotation_axis, rotation_centre
        self.base = instances[0]
        self.mesh = hb.replicate(self.base.mesh, self.n_instants)
        self.schemes, self.controls = self.base.schemes, self.base.controls
        self.p = np.concatenate([c.p for c in instances])
        self.U = np.concatenate([c.U for c in instances])
        self.T = np.concatenate([c.T for c in instances])
        self.omegas = None        # omega list of the (single) HB zone, needed for phase-lag patches
        self.phase_lag = []       # [(owner patch, neighbour patch, IBPA)]: phaseLagCyclic pairs (cyclic pairs of the base mesh)

    def with_phase_lag(self, owner_patch, neighbour_patch, ibpa, omegas):
        """Make the cyclic pair (owner_patch, neighbour_patch) a phaseLagCyclic pair with inter-blade phase angle `ibpa`."""
        self.omegas = np.asarray(omegas, float)
        self.phase_lag.append((owner_patch, neighbour_patch, float(ibpa)))
        return self

    def phase_lag_operators(self):
        """[(patch index in the base mesh, D_pl)] for both sides of every pair (owner +IBPA, neighbour -IBPA)."""
        from . import hb
        out = []
        for a, b, ibpa in self.phase_lag:
            out.append((self.base.mesh.patch_index(a), hb.phase_lag_operator(self.snapshots, self.omegas, ibpa)))
            out.append((self.base.mesh.patch_index(b), hb.phase_lag_operator(self.snapshots, self.omegas, -ibpa)))
        return out

    def partition(self, n_parts, mode="x"):
        """One HBCase per rank: every instance case restricted to the same spatial decomposition (HB instants are not sharded
        across ranks — SURVEY §8e — each rank holds all instances of its cells)."""
        part, meshes = self.base.partition(n_parts, mode)
        out = []
        for r, m in enumerate(meshes):
            insts = []
            for c in self.instances:
                g = m.cell_global
                insts.append(Case(c.name, m, c.R, c.Cp, c.schemes, c.controls, {k: dict(v) for k, v in c.bcs.items()}, c.p[g], c.U[g], c.T[g], mu=c.mu, Pr=c.Pr))
            zone = None if self.zone_of_cell is None else self.zone_of_cell[m.cell_global]
            out.append(HBCase(self.name, insts, self.snapshots, self.D, zone, self.cyl_coords, self.rotation_axis, self.rotation_centre))
        return out

    def apply(self, api):
        b = self.base
        np0 = len(b.mesh.patches)
        for patch, Dpl in self.phase_lag_operators():     # before mesh_set: phase-lag pairs need local halo slots
            for K in range(self.n_instants):
                api.phaselag_set(K * np0 + patch, Dpl[K])
        api.mesh_set(self.mesh)
        api.thermo_set(b.R, b.Cp, b.mu, b.Pr)
        api.schemes_set(self.schemes)
        fid = {"p": capi.FIELD_P, "U": capi.FIELD_U, "T": capi.FIELD_T}
        for K, inst in enumerate(self.instances):
            for patch, fields in inst.bcs.items():
                for field, (kind, params) in fields.items():
                    api.bc_set(f"{patch}@{K}", fid[field], kind, params)
        api.hb_set(self.n_instants, self.D, self.zone_of_cell, self.cyl_coords, self.rotation_axis, self.rotation_centre)
        # per-instance MRF / transport fields in the instance-major layout of the replicated mesh:
        # faces = [internal faces of instance 0, 1, ...][boundary faces of instance 0, 1, ...]
        F = b.mesh.n_internal_faces
        if any(c.mrf is not None for c in self.instances):
            fv, om = zip(*[c.mrf_fields(b.mesh) if c.mrf is not None else (np.zeros(b.mesh.n_faces), np.zeros((b.mesh.n_cells, 3))) for c in self.instances])
            api.mrf_set(np.concatenate([x[:F] for x in fv] + [x[F:] for x in fv]), np.concatenate(om))
        if any(c.transport is not None for c in self.instances):
            tr = [c.transport_fields(b.mesh) for c in self.instances]
            api.transport_set(*[np.concatenate([t[k] for t in tr]) for k in range(4)])
        api.state_set(self.p, self.U, self.T)
        return api


def hb_box(n=6, n_instants=3, omega=2 * np.pi * 40.0, flux="ROE", limiter="vanLeer", cyl=False, zoned=False, seed=0, co=5.0,
           cyclic=True, mu=0.0):
    """Small 3-D box with a cyclic pair, a wall, a symmetry plane and an inlet whose total state oscillates at `omega`:
    n_instants = 2*harmonics+1 time instances coupled by the HB operator D (one frequency, uniform snapshots over one
    period, `selectedPeriod` = 2 pi / omega).  `zoned`: only the cells with x > 0.4 belong to the HB zone (cellZone
    instead of `allMesh`); `cyl`: cylindrical momentum source about the z axis through (0.5, -2, 0)."""
    from . import hb
    harmonics = (n_instants - 1) // 2
    omegas = hb.omega_list([omega], [harmonics])
    snaps, Ds = hb.set_instants([omegas], n_instants, selected_period=2 * np.pi / omega)
    insts = []
    for K, t in enumerate(snaps):
        c = periodic_box(n, flux, limiter, seed + 17 * K, cyclic=cyclic, mu=mu)
        s = np.sin(omega * t)
        c.bcs["ymax"] = {"p": ("fixedValue", (1.05e5 * (1 + 0.03 * s),)), "U": ("inletOutlet", (50.0 + 15.0 * s, 10.0, 0.0)),
                         "T": ("inletOutlet", (310.0 + 4.0 * s,))}
        c.schemes = capi.default_schemes(flux_scheme=flux, limiter_rho=limiter, limiter_U=limiter, limiter_T=limiter,
                                         ddt_scheme="steadyState", pseudo_co_num=co, pseudo_co_num_max=50.0)
        insts.append(c)
    base_mesh = insts[0].mesh
    for c in insts[1:]:
        c.mesh = base_mesh
    zone = None
    if zoned:
        zone = np.where(base_mesh.C[:, 0] > 0.4, 0, -1).astype(np.int32)
    kw = {}
    if cyl:
        kw = dict(cyl_coords=[1], rotation_axis=[0.0, 0.0, 2.0], rotation_centre=[0.5, -2.0, 0.0])
    return HBCase("hb-box", insts, snaps, Ds[0], zone, **kw)


def vki_hb(polymesh_dir, n_instants=3, omega=2 * np.pi * 2000.0, amp=0.02, mu=1.5e-5, co=10.0):
    """C5 (ii): Harmonic Balance on the shipped VKI-LS89 mesh (dbnsFullyImplicitHBFoam, `allMesh` zone, one frequency,
    n_instants = 2*harmonics+1 instances over one period): the inlet total pressure oscillates by +-amp at `omega`
    (an incoming wake / potential disturbance stand-in; the tutorial itself ships no HBProperties)."""
    from . import hb
    harmonics = (n_instants - 1) // 2
    omegas = hb.omega_list([omega], [harmonics])
    snaps, Ds = hb.set_instants([omegas], n_instants, selected_period=2 * np.pi / omega)
    insts = []
    for K, t in enumerate(snaps):
        c = vki_ls89(polymesh_dir, mu=mu, co=co) if K == 0 else None
        if c is None:
            b = insts[0]
            c = Case(b.name, b.mesh, b.R, b.Cp, b.schemes, b.controls, {k: dict(v) for k, v in b.bcs.items()}, b.p, b.U, b.T, mu=b.mu, Pr=b.Pr)
        c.bcs["inlet"] = dict(c.bcs["inlet"], p=("totalPressure", (160500.0 * (1 + amp * np.sin(omega * t)), 1.4)))
        insts.append(c)
    return HBCase("vki-HB", insts, snaps, Ds[0])
