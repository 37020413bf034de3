"""Mesh hierarchy generation (adjacency A, down-sampling D, up-sampling U per level) without psbody.

The reference builds its hierarchy at start-up with `mesh_sampling.generate_transform_matrices(reference_mesh,
ds_factors)` (lib/mesh_sampling.py:40-263, called from main.py:38), which needs the `psbody.mesh` package (its Mesh
class, its vertex-connectivity helpers and its C++ AABB tree) -- not installable here, and the only reason the
reference ships pre-computed fixtures for one hierarchy (8 conv layers, ds_factor 2).  This module restates the
three steps on numpy / scipy so that `--num_conv_layers 4/6` and other `--ds_factor`s work:

  * `qslim_decimator_transformer` -- quadric-error edge collapses that keep one of the two end points (so D is a 0/1 row
    selection), driven by a binary heap whose entries are renamed IN PLACE after every collapse.  The reference does
    that renaming on a `heapq` list without restoring the heap order, so which edge pops next depends on the physical
    heap layout; `_EdgeHeap` reproduces `heapq`'s sift rules on parallel numpy arrays (the renaming becomes two
    vectorised assignments instead of two Python scans of the queue per collapse), which is what makes the result match
    the reference's own fixtures: on the SMPL template the down-sampling matrices and adjacencies come out IDENTICAL to
    `data/transform_matrices/for_demo` and `ds2` (tests/test_mesh_sampling.py).
  * `setup_deformation_transfer` -- every vertex of the finer mesh is expressed in the vertices of its closest triangle
    of the coarser mesh (closest point by an exact search: k-d tree bound + Ericson's region classification; in a face:
    barycentric through a 3x3 least-squares solve, on an edge: least-squares onto the span of the two end points -- the
    reference's formula, whose rows do not sum to one --, at a vertex: 1).
  * `generate_transform_matrices` -- the loop over the factors.

Nothing here is on the GPU path: it runs once per hierarchy, on the host, before `CAPE(...)` is constructed.
"""
import math

import numpy as np
import scipy.sparse as sp


class TriMesh(object):
    """The two attributes of psbody's Mesh that the hierarchy generation uses: v [V, 3] float64, f [F, 3] int."""

    def __init__(self, v=None, f=None, filename=None):
        if filename is not None:
            v, f = load_obj(filename)
        self.v = np.asarray(v, np.float64)
        self.f = np.asarray(f, np.int64).reshape(-1, 3)


def load_obj(filename):
    """Vertices and triangles of a Wavefront OBJ file (what `Mesh(filename=...)` reads at main.py:15)."""
    v, f = [], []
    with open(filename) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                v.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                idx = [int(x.split("/")[0]) - 1 for x in t[1:]]
                for k in range(1, len(idx) - 1):            # fan-triangulate polygons
                    f.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(v, np.float64), np.asarray(f, np.int64)


def get_vert_connectivity(mesh):
    """Sparse [V, V] matrix, nonzero where two vertices share an edge (psbody's helper of the same name): every
    directed face edge contributes 1 in both directions, so an interior edge of a manifold mesh has the value 2 --
    what the shipped A.npy hold; the normalised Laplacian does not depend on that scale."""
    n = len(mesh.v)
    vpv = sp.csc_matrix((n, n))
    for i in range(3):
        a, b = mesh.f[:, i], mesh.f[:, (i + 1) % 3]
        m = sp.csc_matrix((np.ones(len(a)), (a, b)), shape=(n, n))
        vpv = vpv + m + m.T
    return vpv


def get_vertices_per_edge(mesh):
    """[E, 2] vertex pairs, each undirected edge once with the smaller index first."""
    vc = sp.coo_matrix(get_vert_connectivity(mesh))
    e = np.stack([vc.row, vc.col], 1)
    return e[e[:, 0] < e[:, 1]]


def vertex_quadrics(mesh):
    """[V, 4, 4]: per vertex, the sum over its faces of the outer product of the face's normalised plane equation
    (lib/mesh_sampling.py:40-65; the plane is the null vector of [v 1], by SVD as there)."""
    q = np.zeros((len(mesh.v), 4, 4))
    ones = np.ones((3, 1))
    for tri in mesh.f:
        _, _, vt = np.linalg.svd(np.hstack((mesh.v[tri], ones)))
        eq = vt[-1, :].reshape(-1, 1)
        eq = eq / np.linalg.norm(eq[0:3])
        outer = np.outer(eq, eq)
        for k in range(3):
            q[tri[k]] += outer
    return q


class _EdgeHeap(object):
    """Min-heap of (cost, (r, c)) entries with the sift rules of Python's heapq, stored in parallel arrays so that the
    end points can be renamed in place by vectorised assignments (which, as in the reference, does NOT restore the heap
    order: the entry order afterwards is whatever heapq's would be)."""

    def __init__(self, capacity):
        self.cost = np.zeros(capacity)
        self.r = np.zeros(capacity, np.int64)
        self.c = np.zeros(capacity, np.int64)
        self.n = 0

    def _less(self, a, b):
        # tuple order of (cost, (r, c))
        ca, cb = self.cost[a], self.cost[b]
        if ca != cb:
            return ca < cb
        if self.r[a] != self.r[b]:
            return self.r[a] < self.r[b]
        return self.c[a] < self.c[b]

    def _item_less(self, item, b):
        if item[0] != self.cost[b]:
            return item[0] < self.cost[b]
        if item[1] != self.r[b]:
            return item[1] < self.r[b]
        return item[2] < self.c[b]

    def _set(self, pos, item):
        self.cost[pos], self.r[pos], self.c[pos] = item

    def _move(self, dst, src):
        self.cost[dst], self.r[dst], self.c[dst] = self.cost[src], self.r[src], self.c[src]

    def _siftdown(self, startpos, pos, item):
        while pos > startpos:
            parent = (pos - 1) >> 1
            if self._item_less(item, parent):
                self._move(pos, parent)
                pos = parent
                continue
            break
        self._set(pos, item)

    def push(self, cost, r, c):
        if self.n == len(self.cost):
            for name in ("cost", "r", "c"):
                a = getattr(self, name)
                setattr(self, name, np.concatenate([a, np.zeros_like(a)]))
        self.n += 1
        self._siftdown(0, self.n - 1, (cost, r, c))

    def pop(self):
        self.n -= 1
        last = (self.cost[self.n], self.r[self.n], self.c[self.n])
        if self.n == 0:
            return last
        top = (self.cost[0], self.r[0], self.c[0])
        # heapq._siftup: walk the smaller child up to a leaf, drop the former last item there, sift it down
        pos, end = 0, self.n
        child = 1
        while child < end:
            right = child + 1
            if right < end and not self._less(child, right):
                child = right
            self._move(pos, child)
            pos = child
            child = 2 * pos + 1
        self._siftdown(0, pos, last)
        return top

    def rename(self, old, new):
        n = self.n
        self.r[:n][self.r[:n] == old] = new
        self.c[:n][self.c[:n] == old] = new


def _collapse_cost(Qv, r, c, v):
    """lib/mesh_sampling.py:139-151: quadric error of keeping r (destroying c) and of keeping c (destroying r)."""
    Qsum = Qv[r] + Qv[c]
    p1 = np.append(v[r], 1.0).reshape(-1, 1)
    p2 = np.append(v[c], 1.0).reshape(-1, 1)
    destroy_c = float(p1.T.dot(Qsum).dot(p1)[0, 0])
    destroy_r = float(p2.T.dot(Qsum).dot(p2)[0, 0])
    return destroy_c, destroy_r, Qsum


def qslim_decimator_transformer(mesh, factor=None, n_verts_desired=None):
    """Simplify `mesh` to ceil(V * factor) vertices (or n_verts_desired) by quadric-error edge collapses that keep one
    of the two end points.  Returns (new_faces [F', 3], D): D is the sparse 0/1 matrix [V', V] that selects the kept
    vertices, in increasing order of their old index (lib/mesh_sampling.py:113-230)."""
    if factor is None and n_verts_desired is None:
        raise Exception("Need either factor or n_verts_desired.")
    if n_verts_desired is None:
        n_verts_desired = math.ceil(len(mesh.v) * factor)
    Qv = vertex_quadrics(mesh)
    nv = len(mesh.v)
    e = get_vertices_per_edge(mesh)
    adj = sp.csc_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(nv, nv))
    adj = (adj + adj.T).tocoo()
    heap = _EdgeHeap(max(2 * len(e), 16))
    for r, c in zip(adj.row, adj.col):                      # same push order as the reference (column-major COO)
        if r > c:
            continue
        dc, dr, _ = _collapse_cost(Qv, r, c, mesh.v)
        heap.push(dr if dr < dc else dc, int(r), int(c))
    faces = mesh.f.copy()
    nverts_total = nv
    while nverts_total > n_verts_desired:
        cost0, r, c = heap.pop()
        r, c = int(r), int(c)
        if r == c:
            continue
        dc, dr, Qsum = _collapse_cost(Qv, r, c, mesh.v)
        cost = dr if dr < dc else dc
        if cost > cost0:                                    # outdated entry: back with the current cost
            heap.push(cost, r, c)
            continue
        keep, destroy = (r, c) if dc < dr else (c, r)
        faces[faces == destroy] = keep
        heap.rename(destroy, keep)
        Qv[r] = Qsum
        Qv[c] = Qsum
        degenerate = (faces[:, 0] == faces[:, 1]) | (faces[:, 1] == faces[:, 2]) | (faces[:, 2] == faces[:, 0])
        if degenerate.any():
            faces = faces[~degenerate]
        nverts_total = len(np.unique(faces))
    return _get_sparse_transform(faces, nv)


def _get_sparse_transform(faces, num_original_verts):
    """Re-index the surviving vertices 0..V'-1 in increasing order of their old index; D[i, old_i] = 1."""
    left = np.unique(faces)
    new_index = np.zeros(int(faces.max()) + 1, np.int64)
    new_index[left] = np.arange(len(left))
    mtx = sp.csc_matrix((np.ones(len(left)), (np.arange(len(left)), left)), shape=(len(left), num_original_verts))
    return new_index[faces], mtx


def closest_points_on_triangles(p, a, b, c):
    """Closest point of each triangle (a_i, b_i, c_i) to the single point p (Ericson, Real-Time Collision Detection
    5.1.5), with the region it lies in: 0 = inside the face, 1..3 = on edge (a,b) / (b,c) / (c,a), 4..6 = at vertex
    a / b / c -- the `nearest_parts` convention lib/mesh_sampling.py:92-106 decodes."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
    bp = p - b
    d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
    cp = p - c
    d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    n = len(a)
    part = np.zeros(n, np.int64)
    q = np.zeros((n, 3))
    done = np.zeros(n, bool)

    def take(mask, pts, code):
        m = mask & ~done
        q[m] = pts[m]
        part[m] = code
        done[m] = True

    with np.errstate(divide="ignore", invalid="ignore"):
        take((d1 <= 0) & (d2 <= 0), a, 4)
        take((d3 >= 0) & (d4 <= d3), b, 5)
        take((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + (d1 / (d1 - d3))[:, None] * ab, 1)
        take((d6 >= 0) & (d5 <= d6), c, 6)
        take((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + (d2 / (d2 - d6))[:, None] * ac, 3)
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        take((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), b + w[:, None] * (c - b), 2)
        denom = 1.0 / (va + vb + vc)
        take(np.ones(n, bool), a + (vb * denom)[:, None] * ab + (vc * denom)[:, None] * ac, 0)
    return q, part


def nearest_on_mesh(source, points):
    """For every point: (face index, part code, closest point) on the triangle mesh `source` -- what
    `source.compute_aabb_tree().nearest(points, True)` returns in the reference (psbody's C++ AABB tree).  Exact: a
    face can only hold the closest point if its bounding sphere reaches into the ball whose radius is the distance to
    the nearest source VERTEX, so the candidates come from a k-d tree over the face centroids."""
    from scipy.spatial import cKDTree
    v, f = source.v, source.f
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cen = (a + b + c) / 3.0
    rad = np.sqrt(np.maximum(((a - cen) ** 2).sum(1), np.maximum(((b - cen) ** 2).sum(1), ((c - cen) ** 2).sum(1))))
    used = np.unique(f)
    ub = cKDTree(v[used]).query(points)[0]
    tree = cKDTree(cen)
    cands = tree.query_ball_point(points, ub + rad.max() + 1e-12)
    faces = np.zeros(len(points), np.int64)
    parts = np.zeros(len(points), np.int64)
    closest = np.zeros((len(points), 3))
    for i, p in enumerate(points):
        cand = np.asarray(sorted(cands[i]), np.int64)
        near = ((cen[cand] - p) ** 2).sum(1) <= (ub[i] + rad[cand] + 1e-12) ** 2
        cand = cand[near]
        q, part = closest_points_on_triangles(p, a[cand], b[cand], c[cand])
        j = int(np.argmin(((q - p) ** 2).sum(1)))
        faces[i], parts[i], closest[i] = cand[j], part[j], q[j]
    return faces, parts, closest


def setup_deformation_transfer(source, target):
    """Sparse [V_target, V_source] matrix expressing every target vertex in the vertices of its closest source
    triangle (lib/mesh_sampling.py:67-110); three stored entries per row, as in the reference."""
    nt = target.v.shape[0]
    rows = np.repeat(np.arange(nt), 3)
    cols = np.zeros(3 * nt, np.int64)
    coef = np.zeros(3 * nt)
    faces, parts, closest = nearest_on_mesh(source, target.v)
    for i in range(nt):
        tri = source.f[faces[i]]
        cols[3 * i:3 * i + 3] = tri
        n_id = parts[i]
        if n_id == 0:                                   # inside the face: coordinates of the closest point
            A = source.v[tri].T
            coef[3 * i:3 * i + 3] = np.linalg.lstsq(A, closest[i], rcond=None)[0]
        elif n_id <= 3:                                 # on an edge: least squares of the VERTEX onto the two end points
            A = np.vstack((source.v[tri[n_id - 1]], source.v[tri[n_id % 3]])).T
            t = np.linalg.lstsq(A, target.v[i], rcond=None)[0]
            coef[3 * i + n_id - 1] = t[0]
            coef[3 * i + n_id % 3] = t[1]
        else:                                           # at a vertex
            coef[3 * i + n_id - 4] = 1.0
    return sp.csc_matrix((coef, (rows, cols)), shape=(nt, source.v.shape[0]))


def generate_transform_matrices(mesh, factors):
    """(M, A, D, U, E) for down-sampling factors `factors` (main.py:31-38: e.g. [1, 2, 1, 2, 1, 2, 1, 1]): meshes,
    adjacency matrices, down- and up-sampling matrices, edge lists -- lib/mesh_sampling.py:244-263.  `mesh` needs `.v`
    and `.f` (TriMesh, or a psbody Mesh)."""
    mesh = TriMesh(v=mesh.v, f=mesh.f)
    M, A, D, U, E = [mesh], [get_vert_connectivity(mesh)], [], [], [get_vertices_per_edge(mesh)]
    for factor in [1.0 / x for x in factors]:
        ds_f, ds_D = qslim_decimator_transformer(M[-1], factor=factor)
        D.append(ds_D)
        new_mesh = TriMesh(v=ds_D.dot(M[-1].v), f=ds_f)
        M.append(new_mesh)
        A.append(get_vert_connectivity(new_mesh))
        U.append(setup_deformation_transfer(M[-1], M[-2]))
        E.append(get_vertices_per_edge(new_mesh))
    return M, A, D, U, E


def hierarchy(mesh, num_conv_layers=8, ds_factor=2):
    """The operator lists main.py:31-43 hands to `models.CAPE`: (L, D, U, p) for 4, 6 or 8 conv layers."""
    from .topology import laplacian
    if num_conv_layers == 4:
        ds = [1, ds_factor, 1, 1]
    elif num_conv_layers == 6:
        ds = [1, ds_factor, 1, ds_factor, 1, 1]
    elif num_conv_layers == 8:
        ds = [1, ds_factor, 1, ds_factor, 1, ds_factor, 1, 1]
    else:
        raise NotImplementedError("num_conv_layers must be 4, 6 or 8 (main.py:31-36)")
    _, A, D, U, _ = generate_transform_matrices(mesh, ds)
    p = [a.shape[0] for a in A]
    A = [a.astype("float32") for a in A]
    D = [d.astype("float32") for d in D]
    U = [u.astype("float32") for u in U]
    return [laplacian(a, normalized=True) for a in A], D, U, p
