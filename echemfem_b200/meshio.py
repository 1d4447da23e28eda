"""`Mesh("file.msh")` of the reference's scripts (examples/bortels_*.py, `Mesh('bortels_structuredquad_nondim.msh')`).

Two inputs are understood:

1. a gmsh ASCII `.msh` file (format 2.2 or 4.1) of quadrilaterals (2D) or hexahedra (3D): nodes, cells and the
   physical tags of the boundary elements (lines / quads) become a `mesh.Mesh` with the same boundary ids the
   reference passes to `ds(id)`;
2. when the `.msh` file does not exist (the reference ships only the `.geo` sources and gmsh is not available
   offline) but a `.geo` of the same stem does, and it describes a TRANSFINITE rectangle split into segments
   (the `examples/bortels_structuredquad*.geo` family): the structured mesh is built directly from the `.geo`
   -- corner points, `Transfinite Curve{..} = n Using Bump r | Progression r`, `Physical Curve(name, id)` --
   with gmsh's grading laws restated in `meshes.bump_nodes` / `progression_nodes`.

Local vertex order of a cell is lexicographic with the last coordinate fastest (mesh.py); gmsh's is
counter-clockwise (quads) / bottom-then-top counter-clockwise (hexes).
"""
import os
import re

import numpy as np

from .mesh import Mesh, TensorMesh

# gmsh element types -> number of nodes
_NNODES = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8, 15: 1}
_QUAD_TO_LEX = [0, 3, 1, 2]                      # ccw (a, b, c, d) -> (0,0) (0,1) (1,0) (1,1)
_HEX_TO_LEX = [0, 4, 3, 7, 1, 5, 2, 6]           # gmsh hex -> x slowest, z fastest


def read_msh(path, **kw):
    if os.path.exists(path):
        return _from_msh(path)
    geo = os.path.splitext(path)[0] + ".geo"
    if os.path.exists(geo):
        return mesh_from_transfinite_geo(geo)
    raise FileNotFoundError("%s not found (and no %s to build it from)" % (path, geo))


# ------------------------------------------------------------------------------------------ .msh
def _sections(text):
    out = {}
    for m in re.finditer(r"\$(\w+)\n(.*?)\n\$End\1", text, re.S):
        out[m.group(1)] = m.group(2).split("\n")
    return out


def _parse_v2(sec):
    lines = sec["Nodes"]
    n = int(lines[0])
    ids = np.empty(n, dtype=np.int64)
    xyz = np.empty((n, 3))
    for i in range(n):
        t = lines[1 + i].split()
        ids[i] = int(t[0])
        xyz[i] = [float(v) for v in t[1:4]]
    elems = []
    lines = sec["Elements"]
    for i in range(int(lines[0])):
        t = [int(v) for v in lines[1 + i].split()]
        etype, ntags = t[1], t[2]
        phys = t[3] if ntags > 0 else 0
        elems.append((etype, phys, t[3 + ntags:]))
    return ids, xyz, elems


def _parse_v4(sec):
    lines = sec["Nodes"]
    nblocks, n = [int(v) for v in lines[0].split()[:2]]
    ids = np.empty(n, dtype=np.int64)
    xyz = np.empty((n, 3))
    k, pos = 0, 1
    for _ in range(nblocks):
        nb = int(lines[pos].split()[3])
        pos += 1
        for i in range(nb):
            ids[k + i] = int(lines[pos + i])
        pos += nb
        for i in range(nb):
            xyz[k + i] = [float(v) for v in lines[pos + i].split()[:3]]
        pos += nb
        k += nb
    # entity -> physical tag
    ent_phys = {}
    if "Entities" in sec:
        el = sec["Entities"]
        counts = [int(v) for v in el[0].split()]
        pos = 1
        for dim, cnt in enumerate(counts):
            for _ in range(cnt):
                t = el[pos].split()
                pos += 1
                off = 4 if dim == 0 else 7
                nphys = int(t[off])
                if nphys:
                    ent_phys[(dim, int(t[0]))] = int(t[off + 1])
    elems = []
    lines = sec["Elements"]
    nblocks = int(lines[0].split()[0])
    pos = 1
    for _ in range(nblocks):
        dim, tag, etype, nb = [int(v) for v in lines[pos].split()]
        pos += 1
        phys = ent_phys.get((dim, tag), 0)
        for i in range(nb):
            t = [int(v) for v in lines[pos + i].split()]
            elems.append((etype, phys, t[1:]))
        pos += nb
    return ids, xyz, elems


def _from_msh(path):
    with open(path) as f:
        text = f.read().replace("\r\n", "\n")
    sec = _sections(text)
    version = float(sec["MeshFormat"][0].split()[0])
    if int(sec["MeshFormat"][0].split()[1]) != 0:
        raise NotImplementedError("binary .msh files are not supported")
    ids, xyz, elems = _parse_v4(sec) if version >= 4.0 else _parse_v2(sec)
    index = {int(g): i for i, g in enumerate(ids)}
    has_hex = any(e[0] == 5 for e in elems)
    if any(e[0] in (2, 4) for e in elems):
        raise NotImplementedError("only quadrilateral / hexahedral cells are supported")
    dim = 3 if has_hex else 2
    ctype, btype = (5, 3) if dim == 3 else (3, 1)
    order = _HEX_TO_LEX if dim == 3 else _QUAD_TO_LEX
    cells = np.array([[index[v] for v in e[2]] for e in elems if e[0] == ctype], dtype=np.int64)[:, order]
    coords = xyz[:, :dim].copy()
    if dim == 2:
        # clockwise quads: swap the two coordinate directions of the cell
        X = coords[cells]
        e1, e2 = X[:, 2] - X[:, 0], X[:, 1] - X[:, 0]
        neg = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0] < 0
        cells[neg] = cells[neg][:, [0, 2, 1, 3]]
    btag = {}
    for etype, phys, nodes in elems:
        if etype == btype and phys:
            btag[tuple(sorted(index[v] for v in nodes))] = phys

    def marker_of_facet(fcoords, fverts):
        return np.array([btag.get(tuple(sorted(int(v) for v in fv)), 0) for fv in fverts], dtype=np.int64)
    used = np.unique(cells)
    remap = -np.ones(len(coords), dtype=np.int64)
    remap[used] = np.arange(len(used))
    btag = {tuple(sorted(int(remap[v]) for v in k)): p for k, p in btag.items() if all(remap[v] >= 0 for v in k)}
    return Mesh(coords[used], remap[cells], marker_of_facet, name=os.path.splitext(os.path.basename(path))[0])


# ------------------------------------------------------------------------------------------ .geo
def progression_nodes(length, n, r):
    """`Using Progression r`: successive intervals grow by the factor r."""
    if n < 2:
        return np.array([0.0, length])
    if abs(r - 1.0) < 1e-14:
        return np.linspace(0.0, length, n)
    w = r ** np.arange(n - 1)
    x = np.concatenate([[0.0], np.cumsum(w)]) * (length / w.sum())
    x[-1] = length
    return x


def _eval(expr, env):
    return float(eval(expr.strip(), {"__builtins__": {}}, dict(env)))


def mesh_from_transfinite_geo(path):
    """Structured mesh of a `.geo` whose single plane surface is a transfinite rectangle (see module doc)."""
    from .meshes import bump_nodes
    with open(path) as f:
        text = re.sub(r"//[^\n]*", "", f.read())
    env, points, lines = {}, {}, {}
    trans = {}                                        # |curve| -> (n, law, coefficient, reversed)
    phys = {}                                         # physical id -> [curves]
    for stmt in text.split(";"):
        s = stmt.strip()
        if not s:
            continue
        m = re.match(r"^(\w+)\s*=\s*([^{}]+)$", s)
        if m:
            env[m.group(1)] = _eval(m.group(2), env)
            continue
        m = re.match(r"^Point\s*\(\s*(\d+)\s*\)\s*=\s*\{(.*)\}$", s, re.S)
        if m:
            v = [_eval(t, env) for t in m.group(2).split(",")]
            points[int(m.group(1))] = np.array(v[:3])
            continue
        m = re.match(r"^Line\s*\(\s*(\d+)\s*\)\s*=\s*\{(.*)\}$", s, re.S)
        if m:
            a, b = [int(_eval(t, env)) for t in m.group(2).split(",")]
            lines[int(m.group(1))] = (a, b)
            continue
        m = re.match(r"^Transfinite\s+(?:Curve|Line)\s*\{(.*?)\}\s*=\s*([^U]+?)(?:Using\s+(\w+)\s+(.+))?$", s, re.S)
        if m:
            n = int(round(_eval(m.group(2), env)))
            law = m.group(3) or "Progression"
            coef = _eval(m.group(4), env) if m.group(4) else 1.0
            for t in m.group(1).split(","):
                c = int(_eval(t, env))
                trans[abs(c)] = (n, law, coef, c < 0)
            continue
        m = re.match(r"^Physical\s+(?:Curve|Line)\s*\((.*?)\)\s*=\s*\{(.*)\}$", s, re.S)
        if m:
            tag = m.group(1).split(",")[-1]
            phys[int(_eval(tag, env))] = [abs(int(_eval(t, env))) for t in m.group(2).split(",")]
            continue
    if not re.search(r"Transfinite\s+Surface", text) or not re.search(r"Recombine\s+Surface", text):
        raise NotImplementedError("%s: only recombined transfinite surfaces can be built without gmsh" % path)
    P = np.array(list(points.values()))
    lo, hi = P.min(axis=0), P.max(axis=0)
    if np.any(np.abs(P[:, 2]) > 0):
        raise NotImplementedError("%s: planar (z = 0) geometries only" % path)
    tol = 1e-12 * max(1.0, float(np.max(hi - lo)))

    def axis(dirn, level):
        """Node abscissae along direction `dirn` from the boundary curves lying on coordinate `level`."""
        other = 1 - dirn
        segs = []
        for c, (a, b) in lines.items():
            pa, pb = points[a], points[b]
            if abs(pa[other] - level) < tol and abs(pb[other] - level) < tol and abs(pa[dirn] - pb[dirn]) > tol:
                if c not in trans:
                    raise NotImplementedError("%s: curve %d has no Transfinite statement" % (path, c))
                segs.append((min(pa[dirn], pb[dirn]), max(pa[dirn], pb[dirn]), c, pa[dirn] > pb[dirn]))
        segs.sort()
        xs = [np.array([segs[0][0]])]
        for s0, s1, c, backwards in segs:
            n, law, coef, rev = trans[c]
            if law.lower() == "bump":
                t = bump_nodes(s1 - s0, n, coef)
            else:
                t = progression_nodes(s1 - s0, n, coef)
                if backwards != rev:                                  # progression runs along the curve's own direction
                    t = (s1 - s0) - t[::-1]
            xs.append(s0 + t[1:])
        return np.concatenate(xs)
    xs = axis(0, lo[1])
    ys = axis(1, lo[0])
    # physical ids: boundary facets inherit the id of the curve they lie on
    curve_id = {c: pid for pid, cs in phys.items() for c in cs}
    boxes = []
    for c, (a, b) in lines.items():
        if c in curve_id:
            pa, pb = points[a][:2], points[b][:2]
            boxes.append((np.minimum(pa, pb) - tol, np.maximum(pa, pb) + tol, curve_id[c]))

    def marker_of_facet(fcoords, fverts):
        mid = fcoords.mean(axis=1)
        out = np.zeros(fcoords.shape[0], dtype=np.int64)
        for blo, bhi, pid in boxes:
            hit = np.all((mid >= blo) & (mid <= bhi), axis=1)
            out[hit & (out == 0)] = pid
        return out
    return TensorMesh([xs, ys], name=os.path.splitext(os.path.basename(path))[0], marker_of_facet=marker_of_facet)
