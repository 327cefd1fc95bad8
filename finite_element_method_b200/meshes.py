"""Synthetic meshes of the benchmark configurations, as plain numpy arrays.

A mesh is a dict:
  x, y, z            float64[n_nodes]          node index = array position (insertion order)
  t_n1, t_n2         uint32[n_truss]           node indices;  t_E, t_A float64; t_A2 float64|None (NaN = None)
  b_n1, b_n2         uint32[n_beam];           b_props float64[8, n] = E, nu, A, I11, I22, I12, It, ks
                                               b_axis  float64[3, n] = local_axis_1_direction (SoA)
  p_n                uint32[4, n_plate]        reference node order 1:(+,+) 2:(-,+) 3:(-,-) 4:(+,-)
  p_props            float64[4, n]             E, nu, t, ks
  rel_tol, abs_tol, nodes_number
Insertion (= accumulation) order is plates, then beams, then trusses.
"""
from __future__ import annotations

import numpy as np

REL_TOL = 1e-4   # the reference tests' tolerances (tests/fem/test_fem.rs:7-8)
ABS_TOL = 1e-12


def _empty(n_nodes=0):
    return {
        "x": np.zeros(n_nodes), "y": np.zeros(n_nodes), "z": np.zeros(n_nodes),
        "t_n1": np.zeros(0, np.uint32), "t_n2": np.zeros(0, np.uint32),
        "t_E": np.zeros(0), "t_A": np.zeros(0), "t_A2": None,
        "b_n1": np.zeros(0, np.uint32), "b_n2": np.zeros(0, np.uint32),
        "b_props": np.zeros((8, 0)), "b_axis": np.zeros((3, 0)),
        "p_n": np.zeros((4, 0), np.uint32), "p_props": np.zeros((4, 0)),
        "rel_tol": REL_TOL, "abs_tol": ABS_TOL, "nodes_number": n_nodes, "name": "empty",
    }


def n_elements(mesh) -> int:
    return len(mesh["t_n1"]) + len(mesh["b_n1"]) + np.asarray(mesh["p_n"]).reshape(4, -1).shape[1]


def reference_truss_model():
    """The crate's own test model: nodes (0,0,0),(30,0,0), one truss E=1e6, A=2
    (tests/fem/test_fem.rs:10-15)."""
    m = _empty(2)
    m["x"] = np.array([0.0, 30.0])
    m["t_n1"] = np.array([0], np.uint32); m["t_n2"] = np.array([1], np.uint32)
    m["t_E"] = np.array([1e6]); m["t_A"] = np.array([2.0])
    m["name"] = "reference-2-node-truss"
    return m


def _grid3(n, spacing=1.0):
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    idx = (i + n * (j + n * k)).ravel()
    assert np.array_equal(idx, np.arange(n ** 3))
    return (i.ravel() * spacing).astype(np.float64), (j.ravel() * spacing).astype(np.float64), \
           (k.ravel() * spacing).astype(np.float64)


def _lattice_edges(n, diagonals):
    """+x, +y, +z edges (in that order, each in start-node index order), then one body diagonal per cell."""
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    idx = i + n * (j + n * k)
    a, b, kind = [], [], []
    for axis, (di, dj, dk) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
        m = (i + di < n) & (j + dj < n) & (k + dk < n)
        a.append(idx[m]); b.append(idx[m] + di + n * (dj + n * dk)); kind.append(np.full(m.sum(), axis))
    if diagonals:
        m = (i + 1 < n) & (j + 1 < n) & (k + 1 < n)
        a.append(idx[m]); b.append(idx[m] + 1 + n * (1 + n)); kind.append(np.full(m.sum(), 3))
    return np.concatenate(a).astype(np.uint32), np.concatenate(b).astype(np.uint32), np.concatenate(kind)


def truss_cube(n=3, taper_every=5):
    """Config 1(ii): n^3-node cube lattice, axis edges + the 8... body diagonals of every cell,
    E=2.1e11, A=1e-4, every `taper_every`-th element tapered to A2=2e-4. n=3 gives 27 nodes,
    54 axis edges + 8 diagonals = 62 trusses."""
    x, y, z = _grid3(n)
    m = _empty(n ** 3)
    m["x"], m["y"], m["z"] = x, y, z
    a, b, _ = _lattice_edges(n, diagonals=True)
    m["t_n1"], m["t_n2"] = a, b
    ne = len(a)
    m["t_E"] = np.full(ne, 2.1e11); m["t_A"] = np.full(ne, 1e-4)
    a2 = np.full(ne, np.nan); a2[::taper_every] = 2e-4
    m["t_A2"] = a2
    m["name"] = f"truss-cube-{n}"
    return m


def truss_lattice(n=64, n_elements_cap=1_000_000, jitter=False):
    """Config 2 (T): n^3 lattice, +x,+y,+z edges then one body diagonal per cell, first `cap` elements;
    E=2.1e11, A=1e-4*(1+u), u~U[0,1) seed 20240601; jitter adds U(-0.1,0.1) per coordinate (seed 20240602)."""
    x, y, z = _grid3(n)
    if jitter:
        r = np.random.default_rng(20240602)
        x = x + r.uniform(-0.1, 0.1, x.shape); y = y + r.uniform(-0.1, 0.1, y.shape); z = z + r.uniform(-0.1, 0.1, z.shape)
    m = _empty(n ** 3)
    m["x"], m["y"], m["z"] = x, y, z
    a, b, _ = _lattice_edges(n, diagonals=True)
    a, b = a[:n_elements_cap], b[:n_elements_cap]
    ne = len(a)
    u = np.random.default_rng(20240601).random(ne)
    m["t_n1"], m["t_n2"] = a, b
    m["t_E"] = np.full(ne, 2.1e11); m["t_A"] = 1e-4 * (1.0 + u)
    m["name"] = f"truss-lattice-{n}^3-{ne}{'-jitter' if jitter else ''}"
    return m


def beam_frame(n=88, n_elements_cap=2_000_000, jitter=False):
    """Config 3 (B): n^3 grid, +x/+y/+z edges, first `cap`; E=2.1e11, nu=0.3, A=1e-2(1+u), I11=8e-6(1+u),
    I22=4e-6(1+u), I12=0, It=1e-5, ks=5/6 (seed 20240603); axis1=(0,0,1) for x/y members, (1,0,0) for z
    members. Jitter (seed 20240604) perturbs coordinates and uses axis1=(0.1,0.2,1.0)."""
    x, y, z = _grid3(n)
    if jitter:
        r = np.random.default_rng(20240604)
        x = x + r.uniform(-0.1, 0.1, x.shape); y = y + r.uniform(-0.1, 0.1, y.shape); z = z + r.uniform(-0.1, 0.1, z.shape)
    m = _empty(n ** 3)
    m["x"], m["y"], m["z"] = x, y, z
    a, b, kind = _lattice_edges(n, diagonals=False)
    a, b, kind = a[:n_elements_cap], b[:n_elements_cap], kind[:n_elements_cap]
    ne = len(a)
    u = np.random.default_rng(20240603).random(ne)
    m["b_n1"], m["b_n2"] = a, b
    m["b_props"] = np.stack([np.full(ne, 2.1e11), np.full(ne, 0.3), 1e-2 * (1 + u), 8e-6 * (1 + u),
                             4e-6 * (1 + u), np.zeros(ne), np.full(ne, 1e-5), np.full(ne, 5.0 / 6.0)])
    ax = np.zeros((3, ne))
    if jitter:
        ax[0], ax[1], ax[2] = 0.1, 0.2, 1.0
    else:
        ax[2, kind != 2] = 1.0
        ax[0, kind == 2] = 1.0
    m["b_axis"] = ax
    m["name"] = f"beam-frame-{n}^3-{ne}{'-jitter' if jitter else ''}"
    return m


def _plate_conn(nx, ny, j0=0, j1=None):
    """elements (i, j), j in [j0, j1): n1=(i+1,j+1) n2=(i,j+1) n3=(i,j) n4=(i+1,j); node index i+(nx+1)*j."""
    j1 = ny if j1 is None else j1
    jj, ii = np.meshgrid(np.arange(j0, j1), np.arange(nx), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    w = nx + 1
    n3 = ii + w * jj
    return np.stack([n3 + 1 + w, n3 + w, n3, n3 + 1]).astype(np.uint32)


def plate_grid(nx=2000, ny=2000, variant="flat", dx=1.0, dy=0.75, rows=None):
    """Config 4 (P): (nx+1)x(ny+1) nodes at (i*dx, j*dy, 0); E=2.1e11, nu=0.3, t=0.01(1+u), ks=5/6
    (seed 20240605). variant: "flat" | "jitter" (in-plane U(-0.1,0.1)*h, seed 20240606) |
    "x0" (the same mesh in the x=0 plane, Q != I).
    rows=(j0, j1): all nodes, but only the elements of grid rows j0 <= j < j1 — what one rank of a
    row-strip partition owns (identical to local_part() of the whole mesh, without building it)."""
    w, h = nx + 1, ny + 1
    j, i = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    x = (i.ravel() * dx).astype(np.float64); y = (j.ravel() * dy).astype(np.float64); z = np.zeros(w * h)
    if variant == "jitter":
        r = np.random.default_rng(20240606)
        x = x + r.uniform(-0.1, 0.1, x.shape) * dx; y = y + r.uniform(-0.1, 0.1, y.shape) * dy
    elif variant == "x0":
        x, y, z = np.zeros(w * h), x, y  # plane x = 0: (0, i*dx, j*dy); normal is +x
    m = _empty(w * h)
    m["x"], m["y"], m["z"] = x, y, z
    j0, j1 = (0, ny) if rows is None else (max(0, rows[0]), min(ny, rows[1]))
    m["p_n"] = _plate_conn(nx, ny, j0, j1)
    u = np.random.default_rng(20240605).random(nx * ny)[j0 * nx:j1 * nx]
    ne = len(u)
    m["p_props"] = np.stack([np.full(ne, 2.1e11), np.full(ne, 0.3), 0.01 * (1 + u), np.full(ne, 5.0 / 6.0)])
    m["name"] = f"plate-grid-{nx}x{ny}-{variant}"
    return m


def mixed_structure(nx=2000, ny=2000, rows=None, variant="flat"):
    """Config 5 (M): the P node set; nx*ny plates; beams on every +x grid edge of rows j=0..ny-1
    (props as B, axis1=(0,0,1)); trusses on +y grid edges of even columns i=0,2,..,nx-2 (props as T).
    2000x2000 -> 4M plates + 4M beams + 2M trusses = 10M elements. rows, variant: see plate_grid (in the
    "x0" variant the grid lies in the x = 0 plane: beams run along +y with axis1 = (1, 0, 0), trusses along +z)."""
    m = plate_grid(nx, ny, variant, rows=rows)
    w = nx + 1
    j0, j1 = (0, ny) if rows is None else (max(0, rows[0]), min(ny, rows[1]))
    jj, ii = np.meshgrid(np.arange(j0, j1), np.arange(nx), indexing="ij")
    a = (ii + w * jj).ravel().astype(np.uint32)
    nb = len(a)
    u = np.random.default_rng(20240603).random(nx * ny)[j0 * nx:j1 * nx]
    m["b_n1"], m["b_n2"] = a, a + 1
    m["b_props"] = np.stack([np.full(nb, 2.1e11), np.full(nb, 0.3), 1e-2 * (1 + u), 8e-6 * (1 + u),
                             4e-6 * (1 + u), np.zeros(nb), np.full(nb, 1e-5), np.full(nb, 5.0 / 6.0)])
    ax = np.zeros((3, nb)); ax[0 if variant == "x0" else 2] = 1.0
    m["b_axis"] = ax
    jj, ii = np.meshgrid(np.arange(j0, j1), np.arange(0, nx, 2), indexing="ij")
    t = (ii + w * jj).ravel().astype(np.uint32)
    nt = len(t)
    ncol = len(range(0, nx, 2))
    ut = np.random.default_rng(20240601).random(ncol * ny)[j0 * ncol:j1 * ncol]
    m["t_n1"], m["t_n2"] = t, (t + w).astype(np.uint32)
    m["t_E"] = np.full(nt, 2.1e11); m["t_A"] = 1e-4 * (1 + ut)
    m["name"] = f"mixed-{nx}x{ny}" + ("" if variant == "flat" else f"-{variant}")
    return m


def algorithmic_bytes(mesh) -> dict:
    """SURVEY.md §8(d) byte model of one numeric pass: every element record read once
    (truss 24 B, beam 96 B, plate 48 B), every node's coordinates once (24 B), every CSR value of the
    structural block pattern written once (8 B). Indices and maps are symbolic-pass products and are
    not counted."""
    nt, nb = len(mesh["t_n1"]), len(mesh["b_n1"])
    pn = np.asarray(mesh["p_n"]).reshape(4, -1)
    npl = pn.shape[1]
    n_nodes = len(mesh["x"])
    pairs_full, pairs_truss = [], []
    if npl:
        for a in range(4):
            for b in range(4):
                pairs_full.append(pn[a].astype(np.uint64) * n_nodes + pn[b])
    if nb:
        b1, b2 = mesh["b_n1"].astype(np.uint64), mesh["b_n2"].astype(np.uint64)
        pairs_full += [b1 * n_nodes + b1, b1 * n_nodes + b2, b2 * n_nodes + b1, b2 * n_nodes + b2]
    if nt:
        t1, t2 = mesh["t_n1"].astype(np.uint64), mesh["t_n2"].astype(np.uint64)
        pairs_truss += [t1 * n_nodes + t1, t1 * n_nodes + t2, t2 * n_nodes + t1, t2 * n_nodes + t2]
    full = np.unique(np.concatenate(pairs_full)) if pairs_full else np.zeros(0, np.uint64)
    tr = np.unique(np.concatenate(pairs_truss)) if pairs_truss else np.zeros(0, np.uint64)
    tr_only = np.setdiff1d(tr, full, assume_unique=True)
    nnz = 36 * len(full) + 9 * len(tr_only)
    elem = 24 * nt + 96 * nb + 48 * npl
    total = elem + 24 * n_nodes + 8 * nnz
    return {"nnz": int(nnz), "element_bytes": int(elem), "node_bytes": int(24 * n_nodes),
            "value_bytes": int(8 * nnz), "total_bytes": int(total),
            "bytes_per_element": total / max(1, nt + nb + npl)}


def grid_nnz_fast(nx, ny, mixed=False):
    """Closed-form structural nnz of the plate grid (and of M, whose beam/truss pairs are a subset
    of the plate pairs): every node pair sharing an element gets a 6x6 block."""
    w, h = nx + 1, ny + 1
    # node-pair blocks = sum over nodes of (neighbours in the 3x3 stencil inside the grid)
    blocks = (3 * w - 2) * (3 * h - 2)
    return 36 * blocks


def partition_rows(mesh, world: int, grid_width=None):
    """Contiguous node-index ranges per rank (whole grid lines when grid_width is given)."""
    n = len(mesh["x"])
    if grid_width:
        lines = n // grid_width
        cuts = [((lines * r) // world) * grid_width for r in range(world)] + [n]
    else:
        cuts = [(n * r) // world for r in range(world)] + [n]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def local_part(mesh, begin: int, end: int):
    """Elements whose lowest-index node lies in [begin, end) — the rank that owns that node owns the
    element. Node arrays are kept whole."""
    out = dict(mesh)
    pn = np.asarray(mesh["p_n"], np.uint32).reshape(4, -1)
    if pn.shape[1]:
        lo = pn.min(axis=0)
        k = (lo >= begin) & (lo < end)
        out["p_n"] = np.ascontiguousarray(pn[:, k]); out["p_props"] = np.ascontiguousarray(np.asarray(mesh["p_props"]).reshape(4, -1)[:, k])
    if len(mesh["b_n1"]):
        lo = np.minimum(mesh["b_n1"], mesh["b_n2"])
        k = (lo >= begin) & (lo < end)
        out["b_n1"], out["b_n2"] = mesh["b_n1"][k], mesh["b_n2"][k]
        out["b_props"] = np.ascontiguousarray(np.asarray(mesh["b_props"]).reshape(8, -1)[:, k])
        out["b_axis"] = np.ascontiguousarray(np.asarray(mesh["b_axis"]).reshape(3, -1)[:, k])
    if len(mesh["t_n1"]):
        lo = np.minimum(mesh["t_n1"], mesh["t_n2"])
        k = (lo >= begin) & (lo < end)
        out["t_n1"], out["t_n2"] = mesh["t_n1"][k], mesh["t_n2"][k]
        out["t_E"], out["t_A"] = mesh["t_E"][k], mesh["t_A"][k]
        if mesh.get("t_A2") is not None:
            out["t_A2"] = mesh["t_A2"][k]
    return out


def with_node_window(part, begin: int, end: int):
    """A rank's part (local_part / rows=...) reduced to its node window: the nodes it owns, [begin, end), plus the halo
    its elements touch — one contiguous index range [w0, w1). x / y / z are cut to it, `node_window_begin` = w0;
    element connectivity keeps global node indices and `nodes_number` the size of the whole model."""
    out = dict(part)
    hi = [end]
    pn = np.asarray(part["p_n"]).reshape(4, -1)
    if pn.shape[1]:
        hi.append(int(pn.max()) + 1)
    for k in ("b_n1", "b_n2", "t_n1", "t_n2"):
        if len(part[k]):
            hi.append(int(np.max(part[k])) + 1)
    w0, w1 = int(begin), int(max(hi))
    out["x"], out["y"], out["z"] = part["x"][w0:w1], part["y"][w0:w1], part["z"][w0:w1]
    out["node_window_begin"] = w0
    out["nodes_number"] = part.get("nodes_number", len(part["x"]))
    return out


def hub_star(n_spokes=700, beams_every=7):
    """One hub node joined to `n_spokes` rim nodes on a sphere by trusses (every `beams_every`-th spoke
    also carries a beam). The hub's rows are far larger than a slab image, so this mesh drives the
    unstaged assembly kernel; the rim nodes go through the staged one."""
    rng = np.random.default_rng(20240611)
    n = n_spokes + 1
    m = _empty(n)
    v = rng.normal(size=(3, n_spokes))
    v /= np.linalg.norm(v, axis=0)
    v *= 1.0 + rng.uniform(0.0, 0.5, n_spokes)
    v[0] = np.abs(v[0]) + 0.05           # every spoke has a +x component (see the beam note in DESIGN.md)
    m["x"][1:], m["y"][1:], m["z"][1:] = v[0], v[1], v[2]
    rim = np.arange(1, n, dtype=np.uint32)
    m["t_n1"] = np.zeros(n_spokes, np.uint32); m["t_n2"] = rim
    m["t_E"] = np.full(n_spokes, 2.1e11); m["t_A"] = 1e-4 * (1.0 + rng.uniform(0, 1, n_spokes))
    bsel = rim[::beams_every]
    nb = len(bsel)
    u = rng.uniform(0, 1, nb)
    m["b_n1"] = np.zeros(nb, np.uint32); m["b_n2"] = bsel.astype(np.uint32)
    m["b_props"] = np.stack([np.full(nb, 2.1e11), np.full(nb, 0.3), 1e-2 * (1 + u), 8e-6 * (1 + u), 4e-6 * (1 + u),
                             np.zeros(nb), np.full(nb, 1e-5), np.full(nb, 5.0 / 6.0)])
    m["b_axis"] = np.tile(np.array([[0.1], [0.2], [1.0]]), (1, nb))
    m["name"] = f"hub-star-{n_spokes}"
    return m


def folded_plate(nx=12, ny=8):
    """A plate strip folded along a grid line: rows j < ny/2 lie in the z = 0 plane (Q == I, the
    flat fast path), the rest in a plane tilted about the x axis (general Q) — slabs along the fold
    hold both kinds."""
    m = plate_grid(nx, ny, "flat")
    w = nx + 1
    j = np.arange(len(m["x"])) // w
    fold = ny // 2
    y0 = m["y"][fold * w]
    up = j > fold
    dy = m["y"][up] - y0
    m["y"] = m["y"].copy(); m["z"] = m["z"].copy()
    m["y"][up] = y0 + dy * 0.8
    m["z"][up] = dy * 0.6
    m["name"] = f"folded-plate-{nx}x{ny}"
    return m
