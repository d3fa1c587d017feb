"""Synthetic meshes, partitions and model states for the hot-path benchmarks and parity tests.

Definitions follow SURVEY.md section 8(d) / BASELINE.md section 4: square [0,L]^2, nx*nx quads split
into two counter-clockwise triangles with alternating diagonal, interior nodes jittered by
U(-0.2,0.2)*h with numpy.random.default_rng(20240611); ``lat`` linear 70..88 N in y.
Nothing here is on the timed path; it only produces the host arrays a neXtSIM host would own
(mesh, M_conc, M_thick, M_wind, ...) before calling the C ABI.
"""
from dataclasses import dataclass, field

import numpy as np

SEED = 20240611

SIZES = {           # name -> (nx, h [m])
    "toy": (32, 10e3),
    "10km": (316, 10e3),
    "3km": (1000, 3e3),
    "1km": (3162, 1e3),
}


@dataclass
class GlobalMesh:
    nx: int
    h: float
    x: np.ndarray            # [nn] float64
    y: np.ndarray
    tri: np.ndarray          # [ne,3] int32, 1-based file node ids, CCW
    lat: np.ndarray          # [nn] degrees north
    dirichlet_flags_root: np.ndarray   # 1-based node ids on "coast" edges (FE.cpp:323-333)
    neumann_flags_root: np.ndarray     # 1-based node ids on open-boundary edges
    resolution: float = 0.0  # FiniteElement::resolution, FE.cpp:1846-1860

    @property
    def nn(self):
        return self.x.size

    @property
    def ne(self):
        return self.tri.shape[0]


def make_mesh(nx, h, seed=SEED, jitter=0.2, open_east=False):
    """Structured-split triangular mesh with seeded interior jitter."""
    rng = np.random.default_rng(seed)
    n1 = nx + 1
    ii, jj = np.meshgrid(np.arange(n1), np.arange(n1), indexing="xy")   # ii fast (row-major ids)
    x = (ii * h).astype(np.float64).ravel()
    y = (jj * h).astype(np.float64).ravel()
    interior = ((ii > 0) & (ii < nx) & (jj > 0) & (jj < nx)).ravel()
    dx = rng.uniform(-jitter, jitter, size=x.size) * h
    dy = rng.uniform(-jitter, jitter, size=x.size) * h
    x = np.where(interior, x + dx, x)
    y = np.where(interior, y + dy, y)

    qi, qj = np.meshgrid(np.arange(nx), np.arange(nx), indexing="xy")
    qi = qi.ravel()
    qj = qj.ravel()
    n00 = qj * n1 + qi
    n10 = n00 + 1
    n01 = n00 + n1
    n11 = n01 + 1
    even = ((qi + qj) % 2) == 0
    # even quads: diagonal 00-11 ; odd quads: diagonal 10-01 ; both CCW
    t0 = np.where(even[:, None], np.stack([n00, n10, n11], 1), np.stack([n00, n10, n01], 1))
    t1 = np.where(even[:, None], np.stack([n00, n11, n01], 1), np.stack([n10, n11, n01], 1))
    tri = np.empty((2 * nx * nx, 3), np.int64)
    tri[0::2] = t0
    tri[1::2] = t1
    tri = (tri + 1).astype(np.int32)

    L = nx * h
    lat = 70.0 + 18.0 * (y / L)

    ids = np.arange(n1 * n1).reshape(n1, n1)      # [j,i]
    south, north, west, east = ids[0, :], ids[-1, :], ids[:, 0], ids[:, -1]
    if open_east:
        # open boundary: the edges of the east side; their end nodes (corners) also sit on coast
        # edges, and the reference takes neumann = boundary \ dirichlet (FE.cpp:3877-3881)
        dirichlet = np.unique(np.concatenate([south, north, west]))
        neumann = np.setdiff1d(east, dirichlet)
    else:
        dirichlet = np.unique(np.concatenate([south, north, west, east]))
        neumann = np.zeros(0, np.int64)

    m = GlobalMesh(nx=nx, h=h, x=x, y=y, tri=tri, lat=lat,
                   dirichlet_flags_root=(dirichlet + 1).astype(np.int32),
                   neumann_flags_root=(neumann + 1).astype(np.int32))
    m.resolution = float(np.sqrt(element_area(x, y, tri).mean()))
    return m


def named_mesh(name, **kw):
    nx, h = SIZES[name]
    return make_mesh(nx, h, **kw)


def element_area(x, y, tri1):
    a, b, c = tri1[:, 0] - 1, tri1[:, 1] - 1, tri1[:, 2] - 1
    jac = (x[b] - x[a]) * (y[c] - y[a]) - (x[c] - x[a]) * (y[b] - y[a])
    return 0.5 * np.abs(jac)


def element_centroids(m):
    t = m.tri - 1
    return m.x[t].mean(1), m.y[t].mean(1)


# ----------------------------------------------------------------------------------------------
# model state + forcing (host-side vectors named after the FiniteElement members)
# ----------------------------------------------------------------------------------------------
def minstd_uniform01(n):
    """boost::minstd_rand (a=48271, m=2^31-1, seed 1) through boost::uniform_01 (FE.cpp:11464-11468)."""
    out = np.empty(n, np.float64)
    x = 1
    for i in range(n):
        x = (x * 48271) % 2147483647
        out[i] = (x - 1) * (1.0 / 2147483646.0)
    return out


def minstd_uniform01_fast(n):
    """Vectorised equivalent of :func:`minstd_uniform01` (jump-ahead by powers of the multiplier)."""
    m = 2147483647
    out = np.empty(n, np.uint64)
    blk = 1 << 16
    first = np.empty(min(blk, n), np.uint64)
    x = 1
    for i in range(first.size):
        x = (x * 48271) % m
        first[i] = x
    out[:first.size] = first
    jump = pow(48271, blk, m)
    pos = first.size
    cur = first
    while pos < n:
        cur = (cur * np.uint64(jump)) % np.uint64(m)      # < 2^31 * 2^31 fits in uint64
        k = min(blk, n - pos)
        out[pos:pos + k] = cur[:k]
        pos += k
    return (out.astype(np.float64) - 1.0) * (1.0 / 2147483646.0)


def make_state(m, kind="large", seed=SEED, young=True):
    """Global element/node fields.  kind: 'toy' (nextsim.toy.cfg semantics), 'large' (BASELINE.md section 4:
    conc~U(0.85,1), damage~U(0,0.8)) or 'stable' (conc~U(0.95,1), damage~U(0,0.5)).  The 'large' state holds
    very weak elements (viscous relaxation time << sub-cycle) in which the reference's BBM update runs into a
    growing period-2 oscillation of sigma_n around -Pmax; a 1e-15 input perturbation then grows to 1e-7 within
    one model step in the reference algorithm itself ("Conditioning").  'stable' avoids that regime
    and is what the strict 1e-9 full-step parity tests use."""
    rng = np.random.default_rng(seed + 1)
    ne, nn = m.ne, m.nn
    L = m.nx * m.h
    cx, cy = element_centroids(m)
    S = {}
    z_e = np.zeros(ne)
    if kind == "toy":
        ice = cx >= 0.3 * L                      # constant_partial, FE.cpp:11706-11738
        S["M_conc"] = np.where(ice, 1.0, 0.0)
        S["M_thick"] = np.where(ice, 1.0, 0.0)
        S["M_snow_thick"] = z_e.copy()
        S["M_damage"] = z_e.copy()
        S["M_VT"] = np.zeros(2 * nn)
        S["M_wind"] = np.concatenate([np.full(nn, 20.0), np.zeros(nn)])
        S["M_ocean"] = np.zeros(2 * nn)
        S["M_ssh"] = np.zeros(nn)
        S["M_element_depth"] = np.full(ne, 200.0)
        S["M_conc_young"] = z_e.copy()
        S["M_h_young"] = z_e.copy()
        S["M_hs_young"] = z_e.copy()
        S["M_sigma"] = np.zeros((3, ne))
    else:
        conc = rng.uniform(0.85, 1.0, ne) if kind == "large" else rng.uniform(0.95, 1.0, ne)
        # ~8 % of the elements in smooth ice-free patches
        patch = (np.sin(2 * np.pi * 3.0 * cx / L + 0.7) * np.sin(2 * np.pi * 2.0 * cy / L + 1.9)
                 + 0.35 * np.sin(2 * np.pi * 7.0 * (cx + cy) / L))
        thr = np.quantile(patch, 0.92)
        water = patch > thr
        conc = np.where(water, 0.0, conc)
        thick = conc * rng.uniform(0.5, 3.0, ne)
        snow = conc * rng.uniform(0.0, 0.3, ne)
        S["M_conc"] = conc
        S["M_thick"] = thick
        S["M_snow_thick"] = snow
        S["M_damage"] = np.where(water, 0.0, rng.uniform(0.0, 0.8 if kind == "large" else 0.5, ne))
        xn, yn = m.x / L - 0.5, m.y / L - 0.5
        S["M_VT"] = np.concatenate([-0.6 * yn, 0.6 * xn]) * np.tile(np.exp(-4 * (xn ** 2 + yn ** 2)), 2)
        # translating cyclone 5..20 m/s
        xc, yc = xn - 0.15, yn + 0.1
        r2 = xc ** 2 + yc ** 2
        amp = 5.0 + 15.0 * np.exp(-r2 / 0.08)
        ang = np.arctan2(yc, xc) + 0.5 * np.pi + 0.3
        S["M_wind"] = np.concatenate([amp * np.cos(ang), amp * np.sin(ang)])
        S["M_ocean"] = np.concatenate([0.2 * yn, -0.2 * xn]) * np.tile(np.exp(-3 * (xn ** 2 + yn ** 2)), 2)
        S["M_ssh"] = 0.2 * np.sin(2 * np.pi * m.x / L) * np.cos(2 * np.pi * m.y / L)
        S["M_element_depth"] = 5.0 + 400.0 * (0.5 + 0.5 * np.sin(2 * np.pi * cx / L)) * (cy / L)
        if young:
            cy_ = np.where(water, 0.0, np.minimum(1.0 - conc, rng.uniform(0.0, 0.1, ne)))
            S["M_conc_young"] = cy_
            S["M_h_young"] = cy_ * rng.uniform(0.05, 0.25, ne)
            S["M_hs_young"] = cy_ * rng.uniform(0.0, 0.05, ne)
        else:
            S["M_conc_young"] = z_e.copy()
            S["M_h_young"] = z_e.copy()
            S["M_hs_young"] = z_e.copy()
        S["M_sigma"] = np.zeros((3, ne))
    # Dirichlet nodes carry no velocity (closed coast)
    d = m.dirichlet_flags_root - 1
    S["M_VT"][d] = 0.0
    S["M_VT"][d + nn] = 0.0
    S["M_UM"] = np.zeros(2 * nn)
    S["M_UT"] = np.zeros(2 * nn)
    S["M_thick_myi"] = 0.3 * S["M_thick"]
    S["M_conc_myi"] = 0.3 * S["M_conc"]
    S["M_ridge_ratio"] = np.where(S["M_conc"] > 0, 0.1, 0.0)
    S["M_drag_ui"] = np.full(ne, 0.0020)             # ERA5/ECMWF quad_drag_coef_air, options.cpp:335-336
    S["M_drag_ui_young"] = np.full(ne, 0.0020)
    S["M_time_relaxation_damage"] = np.full(ne, 25.0 * 86400.0)
    S["M_random_number"] = minstd_uniform01_fast(ne)
    return S


NODAL2 = ("M_VT", "M_UM", "M_UT", "M_wind", "M_ocean", "D_tau_a", "D_tau_w", "tau_wi")
NODAL1 = ("M_ssh", "lat")
