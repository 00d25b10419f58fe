"""decomposePar input/output for the multi-GPU path (SURVEY.md §8e "Partitioning": use OpenFOAM's own `decomposePar`
output when present).

* `read_decomposed(case_dir, rank)` reads `processor<rank>/constant/polyMesh` as decomposePar writes it (processor patches
  with `myProcNo` / `neighbProcNo`, `cellProcAddressing`, `faceProcAddressing`) and completes the processor-patch geometry
  the way `processorFvPatch::makeWeights` / `makeDeltaCoeffs` do, with the neighbour rank's cell centres taken from its own
  `processor<q>` directory instead of an MPI exchange.  The result is the `Mesh` `icsb200_mesh_set` takes on that rank.
* `write_polymesh`, `box_polymesh` and `decompose` are the stand-ins used where no OpenFOAM installation exists (this
  image): a small hex-box polyMesh writer and a decomposePar look-alike that writes `processorN` directories from a
  cell -> rank list (faces of a processor patch in ascending global face order, reversed on the neighbour side, original
  patches kept — possibly empty — in front, as decomposePar does).

Input tooling only — never inside a timed region.
"""
import os

import numpy as np

from . import PROCESSOR, Mesh, read_polymesh  # noqa: F401

_HEADER = """FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    location    "{loc}";
    object      {obj};
}}

"""


def _write_list(path, cls, obj, rows, loc="constant/polyMesh"):
    with open(path, "w") as f:
        f.write(_HEADER.format(cls=cls, loc=loc, obj=obj))
        f.write(f"{len(rows)}\n(\n")
        f.write("\n".join(rows))
        f.write("\n)\n")


def write_polymesh(directory, points, faces, owner, neighbour, patches):
    """points (P,3) float; faces: list of point-label lists; owner (nFaces), neighbour (nInternalFaces); patches: list of
    dicts {name, type, start, size, + extra boundary-file entries under "entries"}."""
    os.makedirs(directory, exist_ok=True)
    _write_list(os.path.join(directory, "points"), "vectorField", "points", [f"({repr(float(x))} {repr(float(y))} {repr(float(z))})" for x, y, z in points])
    _write_list(os.path.join(directory, "faces"), "faceList", "faces", [f"{len(fc)}({' '.join(str(int(v)) for v in fc)})" for fc in faces])
    _write_list(os.path.join(directory, "owner"), "labelList", "owner", [str(int(v)) for v in owner])
    _write_list(os.path.join(directory, "neighbour"), "labelList", "neighbour", [str(int(v)) for v in neighbour])
    with open(os.path.join(directory, "boundary"), "w") as f:
        f.write(_HEADER.format(cls="polyBoundaryMesh", loc="constant/polyMesh", obj="boundary"))
        f.write(f"{len(patches)}\n(\n")
        for p in patches:
            f.write(f"    {p['name']}\n    {{\n        type            {p['type']};\n")
            for k, v in p.get("entries", {}).items():
                f.write(f"        {k:<15} {v};\n")
            f.write(f"        nFaces          {int(p['size'])};\n        startFace       {int(p['start'])};\n    }}\n")
        f.write(")\n")


def box_polymesh(nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), warp=0.0,
                 patch_types=("patch",) * 6, patch_names=("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")):
    """A hex box in blockMesh numbering (cells and points with i fastest, internal faces in upper-triangular order).
    `warp` > 0 shears the points smoothly so that the mesh is non-orthogonal.  Returns (points, faces, owner, neighbour,
    patches) for `write_polymesh`."""
    pid = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)
    cid = lambda i, j, k: i + nx * (j + ny * k)
    xi, et, ze = np.meshgrid(np.arange(nx + 1) / nx, np.arange(ny + 1) / ny, np.arange(nz + 1) / nz, indexing="ij")
    X = lo[0] + (hi[0] - lo[0]) * (xi + warp * np.sin(np.pi * et) * np.sin(np.pi * ze) * xi * (1 - xi))
    Y = lo[1] + (hi[1] - lo[1]) * (et + warp * np.sin(np.pi * xi) * et * (1 - et))
    Z = lo[2] + (hi[2] - lo[2]) * (ze + warp * np.sin(np.pi * xi) * np.sin(np.pi * et) * ze * (1 - ze))
    points = np.empty(((nx + 1) * (ny + 1) * (nz + 1), 3))
    for k in range(nz + 1):
        for j in range(ny + 1):
            for i in range(nx + 1):
                points[pid(i, j, k)] = (X[i, j, k], Y[i, j, k], Z[i, j, k])
    xf = lambda i, j, k: [pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)]          # normal +x
    yf = lambda i, j, k: [pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)]          # normal +y
    zf = lambda i, j, k: [pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)]          # normal +z
    rev = lambda fc: [fc[0]] + fc[:0:-1]
    faces, owner, neighbour = [], [], []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = cid(i, j, k)
                if i + 1 < nx:
                    faces.append(xf(i + 1, j, k)); owner.append(c); neighbour.append(cid(i + 1, j, k))
                if j + 1 < ny:
                    faces.append(yf(i, j + 1, k)); owner.append(c); neighbour.append(cid(i, j + 1, k))
                if k + 1 < nz:
                    faces.append(zf(i, j, k + 1)); owner.append(c); neighbour.append(cid(i, j, k + 1))
    sides = [
        [(rev(xf(0, j, k)), cid(0, j, k)) for k in range(nz) for j in range(ny)],
        [(xf(nx, j, k), cid(nx - 1, j, k)) for k in range(nz) for j in range(ny)],
        [(rev(yf(i, 0, k)), cid(i, 0, k)) for k in range(nz) for i in range(nx)],
        [(yf(i, ny, k), cid(i, ny - 1, k)) for k in range(nz) for i in range(nx)],
        [(rev(zf(i, j, 0)), cid(i, j, 0)) for j in range(ny) for i in range(nx)],
        [(zf(i, j, nz), cid(i, j, nz - 1)) for j in range(ny) for i in range(nx)],
    ]
    patches = []
    for name, typ, side in zip(patch_names, patch_types, sides):
        patches.append({"name": name, "type": typ, "start": len(faces), "size": len(side)})
        for fc, c in side:
            faces.append(fc); owner.append(c)
    return points, faces, np.array(owner, np.int32), np.array(neighbour, np.int32), patches


def decompose(case_dir, points, faces, owner, neighbour, patches, part):
    """decomposePar look-alike: writes `case_dir/processor<r>/constant/polyMesh` (+ the *ProcAddressing lists) for every
    rank named in the cell -> rank list `part`."""
    part = np.asarray(part)
    owner, neighbour = np.asarray(owner), np.asarray(neighbour)
    F = len(neighbour)
    rev = lambda fc: [fc[0]] + list(fc[:0:-1])
    for r in sorted(set(part.tolist())):
        cells = np.flatnonzero(part == r)
        g2l = -np.ones(len(part), np.int64)
        g2l[cells] = np.arange(len(cells))
        lf, lo, ln, fpa = [], [], [], []           # local faces, owners, neighbours, faceProcAddressing
        for f in range(F):
            if part[owner[f]] == r and part[neighbour[f]] == r:
                lf.append(list(faces[f])); lo.append(g2l[owner[f]]); ln.append(g2l[neighbour[f]]); fpa.append(f + 1)
        lp, bpa = [], []
        for ip, p in enumerate(patches):
            q = {k: v for k, v in p.items()}
            q["start"] = len(lf)
            for f in range(p["start"], p["start"] + p["size"]):
                if part[owner[f]] == r:
                    lf.append(list(faces[f])); lo.append(g2l[owner[f]]); fpa.append(f + 1)
            q["size"] = len(lf) - q["start"]
            lp.append(q); bpa.append(ip)
        proc = {}
        for f in range(F):
            po, pn = part[owner[f]], part[neighbour[f]]
            if po == r and pn != r:
                proc.setdefault(int(pn), []).append((f, False))
            elif pn == r and po != r:
                proc.setdefault(int(po), []).append((f, True))
        for q in sorted(proc):
            start = len(lf)
            for f, flip in proc[q]:
                lf.append(rev(faces[f]) if flip else list(faces[f]))
                lo.append(g2l[neighbour[f] if flip else owner[f]])
                fpa.append(-(f + 1) if flip else f + 1)
            lp.append({"name": f"procBoundary{r}to{q}", "type": "processor", "start": start, "size": len(lf) - start,
                       "entries": {"inGroups": "1(processor)", "matchTolerance": "0.0001", "transform": "unknown", "myProcNo": r, "neighbProcNo": q}})
            bpa.append(-1)
        used = np.unique(np.concatenate([np.asarray(fc) for fc in lf]))
        p2l = -np.ones(len(points), np.int64)
        p2l[used] = np.arange(len(used))
        d = os.path.join(case_dir, f"processor{r}", "constant", "polyMesh")
        write_polymesh(d, np.asarray(points)[used], [[p2l[v] for v in fc] for fc in lf], lo, ln, lp)
        for name, vals in (("cellProcAddressing", cells), ("faceProcAddressing", fpa), ("boundaryProcAddressing", bpa), ("pointProcAddressing", used)):
            _write_list(os.path.join(d, name), "labelList", name, [str(int(v)) for v in vals])


def _read_labels(path):
    txt = open(path).read()
    txt = txt[txt.index("}", txt.index("FoamFile")) + 1:]
    body = txt[txt.index("(") + 1: txt.rindex(")")]
    return np.array(body.split(), np.int64)


def read_decomposed(case_dir, rank, _cache=None):
    """The sub-mesh of `rank` from decomposePar's `processor<rank>/constant/polyMesh`, processor-patch geometry completed
    from the neighbour directories.  `cell_global` / `face_global` come from cellProcAddressing / faceProcAddressing."""
    cache = {} if _cache is None else _cache

    def load(r):
        if r not in cache:
            cache[r] = read_polymesh(os.path.join(case_dir, f"processor{r}", "constant", "polyMesh"))
        return cache[r]

    sub = load(rank)
    d = os.path.join(case_dir, f"processor{rank}", "constant", "polyMesh")
    if os.path.exists(os.path.join(d, "cellProcAddressing")):
        sub.cell_global = _read_labels(os.path.join(d, "cellProcAddressing")).astype(np.int32)
        sub.face_global = (np.abs(_read_labels(os.path.join(d, "faceProcAddressing"))) - 1).astype(np.int32)
    for p in sub.patches:
        if p["kind"] != PROCESSOR:
            continue
        if p["nbr_rank"] < 0:
            raise RuntimeError(f"meshtools: processor patch {p['name']} has no neighbProcNo")
        other = load(p["nbr_rank"])
        match = [x for x in other.patches if x["kind"] == PROCESSOR and x["nbr_rank"] == rank]
        if len(match) != 1 or match[0]["size"] != p["size"]:
            raise RuntimeError(f"meshtools: processor patch {p['name']} has no unique counterpart on rank {p['nbr_rank']}")
        q = match[0]
        f = np.arange(p["start"], p["start"] + p["size"])
        fb = np.arange(q["start"], q["start"] + q["size"])
        if not np.allclose(sub.Cf[f], other.Cf[fb], rtol=0, atol=1e-9 * (1.0 + np.abs(sub.Cf[f]).max())):
            raise RuntimeError(f"meshtools: faces of {p['name']} and {q['name']} do not coincide")
        c_nbr = other.C[other.owner[fb]]                       # processorFvPatch: neighbour cell centres (an MPI exchange in OpenFOAM)
        nf = sub.Sf[f] / sub.magSf[f, None]
        d_own = np.abs(np.einsum("ij,ij->i", nf, sub.Cf[f] - sub.C[sub.owner[f]]))
        d_nbr = np.abs(np.einsum("ij,ij->i", nf, c_nbr - sub.Cf[f]))
        sub.weights[f] = d_nbr / (d_own + d_nbr)
        dv = c_nbr - sub.C[sub.owner[f]]
        md = np.linalg.norm(dv, axis=1)
        sub.deltaCoeffs[f] = 1.0 / md
        sub.nonOrthDeltaCoeffs[f] = 1.0 / np.maximum(np.einsum("ij,ij->i", nf, dv), 0.05 * md)
    return sub


def n_processors(case_dir):
    n = 0
    while os.path.isdir(os.path.join(case_dir, f"processor{n}")):
        n += 1
    return n
