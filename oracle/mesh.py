"""Oracle (test infrastructure): meshes and H1 lattice numbering.

Restates what the reference obtains from MFEM for its conduction solve:

* ``read_mfem_mesh``  - ``Mesh(file, 1)`` with ``fix_orientation=true``
  (/root/reference/src/jots_driver.cpp:156).  Format
  "MFEM mesh v1.0", geometry 3 = SQUARE, 5 = CUBE (header of
  /root/reference/examples/meshes/Reinert_1D_block.mesh:6-14).
* ``uniform_refine``  - ``Mesh::UniformRefinement`` (jots_driver.cpp:168-171,182-185):
  every quad -> 4, every hex -> 8, boundary faces inherit their attribute.
* ``make_brick``      - synthetic axis-aligned brick with the vertex / element /
  boundary-attribute conventions of examples/meshes/cube_25_L0p01.mesh
  (attr 1=x0, 2=y0, 3=z0, 4=x1, 5=y1, 6=z1).
* ``lattice_numbering`` - global numbering of a tensor lattice of (p+1)^dim nodes per
  element, shared on vertices / edges / faces (``H1_FECollection`` +
  ``ParFiniteElementSpace``, jots_driver.cpp:194-196).  The numbering convention is
  OURS (first appearance in (element, local lexicographic node) order): MFEM's native
  order (vertices, edges, faces, interiors) is not part of the results contract.

Local conventions: local node (i,j,k), x fastest: l = i + D*(j + D*k).  MFEM vertex
order for SQUARE: (0,0),(1,0),(1,1),(0,1); CUBE: bottom face counter-clockwise then top.
"""
from __future__ import annotations

import numpy as np

# corner bits (a,b[,c]) -> MFEM local vertex id
QUAD_CORNER = np.array([[0, 1], [3, 2]])  # [b][a]
HEX_CORNER = np.array([[[0, 1], [3, 2]], [[4, 5], [7, 6]]])  # [c][b][a]


def corner_vertex(dim, bits):
    if dim == 2:
        return int(QUAD_CORNER[bits[1]][bits[0]])
    return int(HEX_CORNER[bits[2]][bits[1]][bits[0]])


class Mesh:
    def __init__(self, dim, verts, elems, elem_attr, bdr, bdr_attr):
        self.dim = int(dim)
        self.verts = np.ascontiguousarray(verts, dtype=np.float64)
        self.elems = np.ascontiguousarray(elems, dtype=np.int64)
        self.elem_attr = np.ascontiguousarray(elem_attr, dtype=np.int64)
        self.bdr = np.ascontiguousarray(bdr, dtype=np.int64)
        self.bdr_attr = np.ascontiguousarray(bdr_attr, dtype=np.int64)
        # structured hint (set by make_brick): (nx, ny, nz, lx, ly, lz)
        self.brick = None

    @property
    def n_elem(self):
        return self.elems.shape[0]

    @property
    def bdr_attributes(self):
        """Sorted unique boundary attributes (MFEM ``Mesh::bdr_attributes``)."""
        return np.unique(self.bdr_attr)

    def element_vertices(self):
        """[n_elem, 2^dim, dim] coordinates in MFEM local vertex order."""
        return self.verts[self.elems]


def _elem_jacobian_at_origin(dim, X):
    """Columns dX/dxi_a at the reference origin corner (exact for affine elements)."""
    J = np.empty((X.shape[0], dim, dim))
    v0 = X[:, corner_vertex(dim, (0,) * dim)]
    for a in range(dim):
        bits = [0] * dim
        bits[a] = 1
        J[:, :, a] = X[:, corner_vertex(dim, tuple(bits))] - v0
    return J


def fix_orientation(mesh: Mesh) -> int:
    """Flip negatively oriented elements (MFEM ``Mesh::CheckElementOrientation(true)``)."""
    X = mesh.element_vertices()
    det = np.linalg.det(_elem_jacobian_at_origin(mesh.dim, X))
    bad = det < 0
    if bad.any():
        if mesh.dim == 2:
            e = mesh.elems[bad]
            e[:, [1, 3]] = e[:, [3, 1]]
            mesh.elems[bad] = e
        else:
            e = mesh.elems[bad]
            e = e[:, [4, 5, 6, 7, 0, 1, 2, 3]]
            mesh.elems[bad] = e
    return int(bad.sum())


def read_mfem_mesh(path: str) -> Mesh:
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    lines = [ln for ln in lines if ln and not ln.startswith("#")]
    if not lines[0].startswith("MFEM mesh v1.0"):
        raise ValueError("not an MFEM mesh v1.0 file: %s" % path)

    def section(name):
        return lines.index(name) + 1

    i = section("dimension")
    dim = int(lines[i])
    i = section("elements")
    ne = int(lines[i])
    rows = [[int(t) for t in lines[i + 1 + j].split()] for j in range(ne)]
    geom = {2: 3, 3: 5}[dim]
    for r in rows:
        if r[1] != geom:
            raise ValueError("only SQUARE (2-D) / CUBE (3-D) elements are supported")
    elem_attr = [r[0] for r in rows]
    elems = [r[2:] for r in rows]
    i = section("boundary")
    nb = int(lines[i])
    rows = [[int(t) for t in lines[i + 1 + j].split()] for j in range(nb)]
    bdr_attr = [r[0] for r in rows]
    bdr = [r[2:] for r in rows]
    i = section("vertices")
    nv = int(lines[i])
    vdim = int(lines[i + 1])
    if vdim != dim:
        raise ValueError("space dimension != mesh dimension is not supported")
    verts = [[float(t) for t in lines[i + 2 + j].split()] for j in range(nv)]
    mesh = Mesh(dim, verts, elems, elem_attr, bdr, bdr_attr)
    fix_orientation(mesh)
    return mesh


def make_brick(nx, ny, nz, lx, ly, lz) -> Mesh:
    """Axis-aligned structured hex brick [0,lx]x[0,ly]x[0,lz], vertices lexicographic
    (x fastest), elements lexicographic, boundary attributes 1=x0 2=y0 3=z0 4=x1 5=y1 6=z1."""
    vx, vy, vz = nx + 1, ny + 1, nz + 1
    xs = np.arange(vx) * (lx / nx)
    ys = np.arange(vy) * (ly / ny)
    zs = np.arange(vz) * (lz / nz)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def vid(i, j, k):
        return i + vx * (j + vy * k)

    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    elems = np.stack([vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k),
                      vid(i, j, k + 1), vid(i + 1, j, k + 1), vid(i + 1, j + 1, k + 1), vid(i, j + 1, k + 1)], axis=1)
    bdr, battr = [], []

    def face(attr, a, b, fn):
        B, A = np.meshgrid(np.arange(b), np.arange(a), indexing="ij")
        A, B = A.ravel(), B.ravel()
        bdr.append(np.stack(fn(A, B), axis=1))
        battr.append(np.full(A.size, attr))

    face(1, ny, nz, lambda j, k: [vid(0, j, k), vid(0, j, k + 1), vid(0, j + 1, k + 1), vid(0, j + 1, k)])
    face(2, nx, nz, lambda i, k: [vid(i, 0, k), vid(i + 1, 0, k), vid(i + 1, 0, k + 1), vid(i, 0, k + 1)])
    face(3, nx, ny, lambda i, j: [vid(i, j, 0), vid(i, j + 1, 0), vid(i + 1, j + 1, 0), vid(i + 1, j, 0)])
    face(4, ny, nz, lambda j, k: [vid(nx, j, k), vid(nx, j + 1, k), vid(nx, j + 1, k + 1), vid(nx, j, k + 1)])
    face(5, nx, nz, lambda i, k: [vid(i, ny, k), vid(i, ny, k + 1), vid(i + 1, ny, k + 1), vid(i + 1, ny, k)])
    face(6, nx, ny, lambda i, j: [vid(i, j, nz), vid(i + 1, j, nz), vid(i + 1, j + 1, nz), vid(i, j + 1, nz)])
    mesh = Mesh(3, verts, elems, np.ones(elems.shape[0]), np.concatenate(bdr), np.concatenate(battr))
    mesh.brick = (nx, ny, nz, float(lx), float(ly), float(lz))
    return mesh


# ----------------------------------------------------------------------------------------------
# lattice numbering
# ----------------------------------------------------------------------------------------------

def _node_keys(elems, dim, p):
    """Orientation independent integer keys [n_elem, (p+1)^dim, 5] of lattice nodes."""
    ne = elems.shape[0]
    D = p + 1
    nloc = D ** dim
    keys = np.zeros((ne, nloc, 5), dtype=np.int64)
    eid = np.arange(ne, dtype=np.int64)
    for l in range(nloc):
        idx = [(l // D ** a) % D for a in range(dim)]
        free = [a for a in range(dim) if 0 < idx[a] < p]
        fixed_bits = [0 if idx[a] == 0 else 1 for a in range(dim)]
        if len(free) == 0:
            v = elems[:, corner_vertex(dim, tuple(fixed_bits))]
            keys[:, l, 0] = 0
            keys[:, l, 1] = v
        elif len(free) == dim:
            keys[:, l, 0] = 3
            keys[:, l, 1] = eid
            keys[:, l, 4] = l
        elif len(free) == 1:
            a = free[0]
            b0 = list(fixed_bits); b0[a] = 0
            b1 = list(fixed_bits); b1[a] = 1
            g0 = elems[:, corner_vertex(dim, tuple(b0))]
            g1 = elems[:, corner_vertex(dim, tuple(b1))]
            t = idx[a]
            fwd = g0 < g1
            keys[:, l, 0] = 1
            keys[:, l, 1] = np.where(fwd, g0, g1)
            keys[:, l, 2] = np.where(fwd, g1, g0)
            keys[:, l, 4] = np.where(fwd, t, p - t)
        else:  # face of a hex: two free axes
            a, b = free
            s, t = idx[a], idx[b]
            c = np.empty((ne, 2, 2), dtype=np.int64)  # [ca][cb]
            for ca in (0, 1):
                for cb in (0, 1):
                    bits = list(fixed_bits); bits[a] = ca; bits[b] = cb
                    c[:, ca, cb] = elems[:, corner_vertex(dim, tuple(bits))]
            flat = c.reshape(ne, 4)
            m = np.argmin(flat, axis=1)
            ca, cb = m // 2, m % 2
            vmin = flat[eid, m]
            na = c[eid, 1 - ca, cb]
            nb = c[eid, ca, 1 - cb]
            sp = np.where(ca == 0, s, p - s)
            tp = np.where(cb == 0, t, p - t)
            a_first = na < nb
            keys[:, l, 0] = 2
            keys[:, l, 1] = vmin
            keys[:, l, 2] = np.where(a_first, na, nb)
            keys[:, l, 3] = np.where(a_first, nb, na)
            keys[:, l, 4] = np.where(a_first, sp * D + tp, tp * D + sp)
    return keys


def lattice_numbering(elems, dim, p):
    """Return (gather [n_elem,(p+1)^dim] int64, n_nodes).  Nodes are numbered in order of
    first appearance when elements are visited in order and local nodes lexicographically."""
    elems = np.asarray(elems, dtype=np.int64)
    keys = _node_keys(elems, dim, p)
    flat = keys.reshape(-1, 5)
    _, first, inv = np.unique(flat, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    gather = rank[inv].reshape(elems.shape[0], -1)
    return gather, int(order.size)


def brick_lattice_numbering(nx, ny, nz, p):
    """Global lexicographic numbering (x fastest) of the H1 lattice of a structured brick."""
    D = p + 1
    NX, NY = nx * p + 1, ny * p + 1
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (i.ravel() * p) + NX * ((j.ravel() * p) + NY * (k.ravel() * p))
    c, b, a = np.meshgrid(np.arange(D), np.arange(D), np.arange(D), indexing="ij")
    off = a.ravel() + NX * (b.ravel() + NY * c.ravel())
    gather = base[:, None] + off[None, :]
    return gather.astype(np.int64), int(NX * NY * (nz * p + 1))


def uniform_refine(mesh: Mesh) -> Mesh:
    """One level of uniform refinement (quad->4, hex->8); geometry is (multi)linear."""
    dim = mesh.dim
    gather, nn = lattice_numbering(mesh.elems, dim, 2)
    # coordinates of the 3^dim lattice (equispaced) by multilinear interpolation
    X = mesh.element_vertices()
    xi = np.array([0.0, 0.5, 1.0])
    nloc = 3 ** dim
    newv = np.zeros((nn, dim))
    for l in range(nloc):
        idx = [(l // 3 ** a) % 3 for a in range(dim)]
        pt = np.zeros((mesh.n_elem, dim))
        for cbits in np.ndindex(*(2,) * dim):
            w = 1.0
            for a in range(dim):
                w *= xi[idx[a]] if cbits[a] else 1.0 - xi[idx[a]]
            if w != 0.0:
                pt += w * X[:, corner_vertex(dim, cbits)]
        newv[gather[:, l]] = pt
    # children
    children = []
    for sub in np.ndindex(*(2,) * dim):  # sub = (sa, sb[, sc]) with index order (a fastest below)
        child = np.empty((mesh.n_elem, 2 ** dim), dtype=np.int64)
        for cbits in np.ndindex(*(2,) * dim):
            idx = [sub[a] + cbits[a] for a in range(dim)]
            l = sum(idx[a] * 3 ** a for a in range(dim))
            child[:, corner_vertex(dim, cbits)] = gather[:, l]
        children.append(child)
    elems = np.stack(children, axis=1).reshape(-1, 2 ** dim)
    elem_attr = np.repeat(mesh.elem_attr, 2 ** dim)
    # boundary faces: locate (element, local face) and split with the same lattice
    be, bf = locate_boundary_faces(mesh)
    nb = mesh.bdr.shape[0]
    if dim == 2:
        newb = np.empty((nb, 2, 2), dtype=np.int64)
        for f in range(4):
            sel = bf == f
            if not sel.any():
                continue
            nodes = face_lattice_nodes(2, 2, f)  # 3 local lattice ids along the edge
            g = gather[be[sel]][:, nodes]
            newb[sel, 0] = g[:, [0, 1]]
            newb[sel, 1] = g[:, [1, 2]]
        bdr = newb.reshape(-1, 2)
        bdr_attr = np.repeat(mesh.bdr_attr, 2)
    else:
        newb = np.empty((nb, 4, 4), dtype=np.int64)
        for f in range(6):
            sel = bf == f
            if not sel.any():
                continue
            nodes = np.asarray(face_lattice_nodes(3, 2, f)).reshape(3, 3)  # [t][s]
            g = gather[be[sel]][:, nodes.ravel()].reshape(-1, 3, 3)
            c = 0
            for tt in (0, 1):
                for ss in (0, 1):
                    newb[sel, c] = np.stack([g[:, tt, ss], g[:, tt, ss + 1], g[:, tt + 1, ss + 1], g[:, tt + 1, ss]], axis=1)
                    c += 1
        bdr = newb.reshape(-1, 4)
        bdr_attr = np.repeat(mesh.bdr_attr, 4)
    out = Mesh(dim, newv, elems, elem_attr, bdr, bdr_attr)
    fix_orientation(out)
    return out


# local face f of the reference element: axis = f // 2, side = f % 2 (0 -> xi=0, 1 -> xi=1)
def face_lattice_nodes(dim, p, f):
    """Local lattice ids of the (p+1)^(dim-1) nodes on local face f, lexicographic in the
    remaining (free) axes, lower axis fastest."""
    D = p + 1
    axis, side = f // 2, f % 2
    free = [a for a in range(dim) if a != axis]
    out = []
    for m in range(D ** (dim - 1)):
        idx = [0] * dim
        idx[axis] = p if side else 0
        mm = m
        for a in free:
            idx[a] = mm % D
            mm //= D
        out.append(sum(idx[a] * D ** a for a in range(dim)))
    return out


def face_vertices_local(dim, f):
    """MFEM-local vertex ids of the corners of local face f."""
    axis, side = f // 2, f % 2
    free = [a for a in range(dim) if a != axis]
    out = []
    for m in range(2 ** (dim - 1)):  # lower free axis fastest
        bits = [0] * dim
        bits[axis] = side
        for n, a in enumerate(free):
            bits[a] = (m >> n) & 1
        out.append(corner_vertex(dim, tuple(bits)))
    return out


def locate_boundary_faces(mesh: Mesh):
    """For every boundary element return (element id, local face id) of the element face
    with the same vertex set."""
    dim = mesh.dim
    nf = 2 * dim
    table = {}
    for f in range(nf):
        lv = face_vertices_local(dim, f)
        fv = np.sort(mesh.elems[:, lv], axis=1)
        for e, key in enumerate(map(tuple, fv)):
            table.setdefault(key, (e, f))
    be = np.empty(mesh.bdr.shape[0], dtype=np.int64)
    bf = np.empty(mesh.bdr.shape[0], dtype=np.int64)
    for b, key in enumerate(map(tuple, np.sort(mesh.bdr, axis=1))):
        be[b], bf[b] = table[key]
    return be, bf
