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


THERMO_FORCING = ("M_tair", "M_mixrat", "M_dair", "M_sphuma", "M_mslp", "M_Qsw_in", "M_Qlw_in", "M_tcc", "M_precip", "M_snowfall",
                  "M_snowfr", "M_mld", "M_ocean_temp", "M_ocean_salt", "M_conc_upd")
THERMO_ICE = ("M_conc", "M_thick", "M_snow_thick", "M_conc_young", "M_h_young", "M_hs_young", "M_ridge_ratio", "M_conc_myi",
              "M_thick_myi", "M_drag_ui", "M_drag_ui_young", "M_time_relaxation_damage")
THERMO_STATE = ("M_sst", "M_sss", "M_tice0", "M_tice1", "M_tice2", "M_tsurf_young", "M_del_vi_tend", "M_freeze_days",
                "M_freeze_onset", "M_conc_summer", "M_thick_summer", "M_fyi_fraction", "M_age_det", "M_age", "M_pond_volume",
                "M_lid_volume", "M_drag_ti", "M_drag_ti_young", "D_pond_fraction")
THERMO_DIAG = ("D_tau_ow", "D_Qa", "D_Qsw", "D_Qlw", "D_Qsh", "D_Qlh", "D_Qo", "D_Qnosun", "D_Qsw_ocean", "D_Qassim", "D_delS",
               "D_fwflux_ice", "D_fwflux", "D_brine", "D_evap", "D_rain", "D_vice_melt", "D_del_vi_young", "D_del_hi",
               "D_del_hi_young", "D_newice", "D_mlt_top", "D_mlt_bot", "D_snow2ice", "D_albedo", "D_sialb", "D_del_ci_mlt_myi",
               "D_del_vi_mlt_myi", "D_del_ci_rplnt_myi", "D_del_vi_rplnt_myi")


def make_thermo_state(ne, nn, seed=SEED, young=True, season="mixed", centroids=None, films=True):
    """Element fields of FiniteElement::thermo() (forcing, ice state, slab ocean, tracers) plus nodal wind / VT / ocean.

    Built to reach every branch of thermo(): ice-free, thin (< hmin after melt), young-only and thick-ice elements; air
    from -35 C (new ice in leads) to +8 C (surface melt, melt ponds); supercooled and warm mixed layers; snow-free and
    snow-covered ice; multi-year-ice tracers on both sides of their clamps.  season: 'winter' / 'summer' bias the air
    temperature and short-wave, 'mixed' spans both.  centroids=(cx, cy, L): every field varies smoothly in space, as
    real fields do (an ice edge, a front), instead of independently element by element -- what the timing runs use,
    since neighbouring elements then mostly take the same branches; the value distributions are the same.
    films=False leaves out the micrometre-thin ice films: the Winton temperature solve of such a film cancels eight digits
    (K32 = 2 ki / hi ~ 1e7), so one ulp of difference becomes 1e-8 -- fine for the bit-for-bit CPU comparison and for single
    calls, not for a 1e-9 comparison of two libm's over several calls."""
    rng = np.random.default_rng(seed + 77)
    if centroids is None:
        u = rng.uniform
        rnd = lambda: rng.random(ne)
        kind = rng.integers(0, 8, ne)                # 0: open water, 1: trace ice, 2: thin ice, 3..7: pack ice
    else:
        # every variate is a smooth random field of the element centroid, rank-transformed so that its marginal is
        # still exactly uniform: same value distribution as above, but spatially coherent
        cx, cy, L = centroids
        xs, ys = 2 * np.pi * cx / L, 2 * np.pi * cy / L

        def rnd(n=ne):
            f = np.zeros(ne)
            for _ in range(4):
                kx, ky = rng.uniform(-3.0, 3.0, 2)
                f += rng.uniform(0.3, 1.0) * np.sin(kx * xs + ky * ys + rng.uniform(0, 2 * np.pi))
            f += 0.02 * rng.standard_normal(ne)
            r = np.empty(ne)
            r[np.argsort(f, kind="stable")] = (np.arange(ne) + 0.5) / ne
            return r

        def u(lo, hi, n=ne):
            if n != ne:
                return rng.uniform(lo, hi, n)        # nodal vectors
            return lo + (hi - lo) * rnd()

        kind = np.minimum((8 * rnd()).astype(np.int64), 7)
    S = {}
    conc = np.select([kind == 0, kind == 1, kind == 2], [0.0, u(1e-13, 0.05, ne), u(0.05, 0.6, ne)], u(0.6, 1.0, ne))
    hice = np.select([kind == 1, kind == 2], [u(0.005, 0.05, ne), u(0.008, 0.4, ne)], u(0.3, 3.5, ne))
    hsnow = np.where(rnd() < 0.3, 0.0, u(0.0, 0.45, ne))
    # rare corners: films of ice thin enough to sublimate or melt away within one step (thermoWinton's layer cascade,
    # FE.cpp:6721-6744, 6760-6764), and cells that are exactly full (lateral melt of melt_type 1, FE.cpp:5577-5580)
    film = (rng.random(ne) < 0.012) & bool(films)
    hice = np.where(film & (conc > 0), 10.0 ** rng.uniform(-7.5, -5.0, ne), hice)
    hsnow = np.where(film, 0.0, hsnow)
    conc = np.where((rng.random(ne) < 0.01) & (conc > 0.6), 1.0, conc)
    S["M_conc"] = conc
    S["M_thick"] = conc * hice
    S["M_snow_thick"] = conc * hsnow
    if young:
        cy = np.minimum(1.0 - conc, np.where(rnd() < 0.35, 0.0, u(0.0, 0.3, ne)))
        S["M_conc_young"] = cy
        S["M_h_young"] = cy * u(0.03, 0.6, ne)       # on both sides of h_young_min (0.05) and h_young_max_sharp (0.275)
        S["M_hs_young"] = cy * np.where(rnd() < 0.4, 0.0, u(0.0, 0.08, ne))
    else:
        S["M_conc_young"] = np.zeros(ne)
        S["M_h_young"] = np.zeros(ne)
        S["M_hs_young"] = np.zeros(ne)
    S["M_ridge_ratio"] = np.where(conc > 0, u(0.0, 0.6, ne), 0.0)
    S["M_conc_myi"] = conc * u(0.0, 1.1, ne)
    S["M_thick_myi"] = S["M_thick"] * u(0.0, 1.1, ne)
    for k in ("M_drag_ui", "M_drag_ui_young", "M_drag_ti", "M_drag_ti_young"):
        S[k] = u(0.8e-3, 3.0e-3, ne)
    S["M_time_relaxation_damage"] = np.full(ne, 25.0 * 86400.0)

    lo, hi = {"winter": (-35.0, -2.0), "summer": (-3.0, 8.0)}.get(season, (-35.0, 8.0))
    tair = u(lo, hi, ne)
    S["M_tair"] = tair
    S["M_dair"] = tair - u(0.0, 6.0, ne)
    S["M_mixrat"] = u(1e-4, 5e-3, ne)
    S["M_sphuma"] = np.where(rnd() < 0.05, -1e-6, u(1e-4, 5e-3, ne))        # round-off negatives are clamped (FE.cpp:4982)
    S["M_mslp"] = u(96500.0, 104500.0, ne)
    S["M_Qsw_in"] = np.where(rnd() < 0.3, 0.0, u(0.0, 120.0 if season == "winter" else 420.0, ne))
    S["M_Qlw_in"] = u(140.0, 340.0, ne)
    S["M_tcc"] = u(0.0, 1.0, ne)
    S["M_precip"] = np.where(rnd() < 0.3, 0.0, u(0.0, 8e-5, ne))              # kg/m^2/s
    S["M_snowfall"] = S["M_precip"] * u(-0.02, 1.0, ne)                                # slight negatives: input round-off (FE.cpp:5343)
    S["M_snowfr"] = u(0.0, 1.0, ne)
    S["M_mld"] = u(5.0, 60.0, ne)
    S["M_ocean_temp"] = u(-1.85, 4.0, ne)
    S["M_ocean_salt"] = u(29.0, 35.5, ne)
    S["M_conc_upd"] = np.where(rnd() < 0.5, 0.0, u(-0.3, 0.1, ne))

    sss = np.where(rnd() < 0.03, u(2.0, 6.0, ne), u(27.0, 35.5, ne))          # a few brackish cells: si_eff = sss < si
    S["M_sss"] = sss
    tf = -0.055 * sss
    S["M_sst"] = np.where(conc + S["M_conc_young"] > 0, tf + u(-0.02, 0.5, ne), tf + u(-0.05, 5.0, ne))
    tfr_ice = -0.055 * 5.0
    S["M_tice0"] = np.where(conc > 0, np.minimum(u(-32.0, 0.0, ne), np.where(hsnow > 0, 0.0, tfr_ice)), tfr_ice)
    S["M_tice1"] = np.where(conc > 0, u(-16.0, -0.4, ne), tfr_ice)
    S["M_tice2"] = np.where(conc > 0, u(-9.0, -0.4, ne), tfr_ice)
    S["M_tsurf_young"] = np.where(S["M_conc_young"] > 0, u(-28.0, tfr_ice, ne), tfr_ice)
    S["M_del_vi_tend"] = u(-0.02, 0.02, ne) * 86400.0
    S["M_freeze_days"] = np.floor(6 * rnd())
    S["M_freeze_onset"] = np.floor(2 * rnd())
    S["M_conc_summer"] = u(0.0, 1.0, ne)
    S["M_thick_summer"] = u(0.0, 3.0, ne)
    S["M_fyi_fraction"] = u(0.0, 1.0, ne)
    S["M_age_det"] = u(0.0, 4e7, ne)
    S["M_age"] = u(0.0, 4e7, ne)
    pond = (rnd() < 0.5) & (conc > 0.1)
    S["D_pond_fraction"] = np.where(pond, u(0.0, 0.35, ne), 0.0)
    S["M_pond_volume"] = np.where(pond, S["D_pond_fraction"] * u(0.0, 0.3, ne), 0.0)
    S["M_lid_volume"] = np.where(pond & (rnd() < 0.5), u(0.0, 0.03, ne), 0.0)

    S["M_wind"] = u(-18.0, 18.0, 2 * nn)
    S["M_VT"] = u(-0.4, 0.4, 2 * nn)
    S["M_ocean"] = u(-0.25, 0.25, 2 * nn)
    return S
