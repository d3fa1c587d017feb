"""A SECOND, independent restatement of the sub-cycled solve, written in vectorised NumPy straight from the reference
text (model/finiteelement.cpp:10182-10575, 4137-4260, 10649-10726, model/constants.hpp:56-86) without looking at
oracle/nextsim_oracle.cpp.  TEST INFRASTRUCTURE ONLY.

The reference ships no fixture for this path, so nothing can pin the C++ oracle bit for bit; what this module buys is
that two restatements written in different styles (scalar loops with the reference's order in C++, whole-array
expressions with np.add.at here) must agree to rounding, which catches transcription slips in either.  Single rank
only (no ghost nodes), no OASIS terms; summation orders differ from the reference, so agreement is to ~1e-12, not
bit-exact.
"""
import numpy as np

RHOI, RHOW, RHOS, RHOA = 917.0, 1025.0, 330.0, 1.22          # constants.hpp:56-86
GRAVITY, OMEGA = 9.80616, 7.292e-5


class SubcycledSolve:
    def __init__(self, x, y, tri1, mask_dirichlet, neumann_flags, lat, p, fields):
        self.p = p
        self.nn, self.ne = x.size, tri1.shape[0]
        self.t = tri1.astype(np.int64) - 1
        self.x, self.y, self.lat = x, y, lat
        self.dir = mask_dirichlet.astype(bool)
        self.neumann = neumann_flags
        f = {k: (np.array(v, float) if k != "M_sigma" else [np.array(s, float) for s in v]) for k, v in fields.items()}
        self.f = f
        self.young_ice = (p.ice_cat_type == 1)
        self.prep()

    # FE.cpp:10234-10414
    def prep(self):
        f, p, t, nn = self.f, self.p, self.t, self.nn
        um = f["M_UM"]
        vx = self.x[t] + um[t]                                   # vertices on the moved mesh
        vy = self.y[t] + um[t + nn]
        side = np.stack([np.hypot(vx[:, 1] - vx[:, 0], vy[:, 1] - vy[:, 0]),
                         np.hypot(vx[:, 2] - vx[:, 1], vy[:, 2] - vy[:, 1]),
                         np.hypot(vx[:, 2] - vx[:, 0], vy[:, 2] - vy[:, 0])], 1)
        # std::accumulate(begin, end, 0) with an INT initial value: truncation after every addition, then int / size_t
        acc = np.zeros(self.ne, np.int64)
        for k in range(3):
            acc = np.trunc(acc + side[:, k]).astype(np.int64)
        self.delta_x = (acc // 3).astype(float)
        jac = (vx[:, 1] - vx[:, 0]) * (vy[:, 2] - vy[:, 0]) - (vx[:, 2] - vx[:, 0]) * (vy[:, 1] - vy[:, 0])
        self.surface = 0.5 * np.abs(jac)
        k1, k2 = (np.arange(3) + 1) % 3, (np.arange(3) + 2) % 3
        self.dxN = (vy[:, k1] - vy[:, k2]) / jac[:, None]
        self.dyN = (vx[:, k2] - vx[:, k1]) / jac[:, None]

        conc, thick, snow = f["M_conc"].copy(), f["M_thick"].copy(), f["M_snow_thick"].copy()
        if self.young_ice:
            conc += f["M_conc_young"]; thick += f["M_h_young"]; snow += f["M_hs_young"]
        emass = np.where(conc > 0, (RHOI * thick + RHOS * snow) / np.where(conc > 0, conc, 1.0), 0.0)
        ssh = f["M_ssh"]
        essh = (ssh[t[:, 0]] + ssh[t[:, 1]] + ssh[t[:, 2]]) / 3.0
        depth_eff = np.maximum(0.0, essh + np.maximum(2.0, f["M_element_depth"]))
        if p.basal_stress_type == 1:
            keel = np.minimum(p.basal_k1 * f["M_thick"], f["M_conc"] * 28.0)
            crit_h = f["M_conc"] * depth_eff / p.basal_k1
            crit_h_mod = keel / p.basal_k1
        else:
            crit_h = crit_h_mod = np.zeros(self.ne)
        ecbu = p.basal_k2 * np.maximum(0.0, crit_h_mod - crit_h) * np.exp(-p.basal_Cb * (1.0 - f["M_conc"]))
        area = np.zeros(nn); mass = np.zeros(nn); cbu = np.zeros(nn)
        for i in range(3):
            np.add.at(area, t[:, i], self.surface)
            np.add.at(mass, t[:, i], emass * self.surface)
            np.maximum.at(cbu, t[:, i], ecbu)
        # grad(m g ssh): skipped on closed boundaries (and, in the running test of the reference, while the node has
        # seen no mass yet -- those contributions are zero anyway)
        mgA3 = emass * self.surface * (GRAVITY / 3.0)
        gs = np.zeros(2 * nn)
        for i in range(3):
            ok = ~self.dir[t[:, i]]
            gx = np.zeros(self.ne); gy = np.zeros(self.ne)
            for j in range(3):
                gx += self.dxN[:, j] * mgA3 * ssh[t[:, j]]
                gy += self.dyN[:, j] * mgA3 * ssh[t[:, j]]
            np.subtract.at(gs, t[ok, i], gx[ok])
            np.subtract.at(gs, t[ok, i] + nn, gy[ok])
        self.grad_ssh = gs
        # nodes
        vt = f["M_VT"]
        ow = (mass == 0.0)
        vt[:nn][ow] = 0.0
        vt[nn:][ow] = 0.0
        dragp = f["M_drag_ui"].copy()
        if self.young_ice:
            c2 = f["M_conc"] + f["M_conc_young"]
            mix = (f["M_drag_ui"] * f["M_conc"] + f["M_drag_ui_young"] * f["M_conc_young"]) / np.where(c2 > 0, c2, 1.0)
            dragp = np.where(c2 > 0, mix, dragp)
        dsum = np.zeros(nn)
        for i in range(3):
            np.add.at(dsum, t[:, i], dragp * self.surface)
        wind = f["M_wind"]
        drag = dsum * RHOA * np.hypot(wind[:nn], wind[nn:]) / area
        self.tau_a = np.concatenate([drag * wind[:nn], drag * wind[nn:]])
        self.fcor = 2 * OMEGA * np.sin(self.lat * np.pi / 180.0)
        rl = 1.0 / area
        self.node_mass = mass * rl
        self.rlmass = 3.0 * rl
        self.cbu = cbu
        self.VTM = vt.copy()

    # FE.cpp:4137-4260
    def update_sigma_damage(self, dt):
        f, p, t, nn = self.f, self.p, self.t, self.nn
        vt = f["M_VT"]
        u, v = vt[t], vt[t + nn]
        e0 = (self.dxN * u).sum(1)
        e1 = (self.dyN * v).sum(1)
        e2 = (self.dyN * u + self.dxN * v).sum(1)
        s = f["M_sigma"]
        d = f["M_damage"]
        conc = f["M_conc"]
        ice = conc > 0.1
        sn = (s[0] + s[1]) * 0.5
        expC = np.exp(p.compaction_param * (1.0 - conc))
        tv = p.undamaged_time_relaxation_sigma * np.power((1.0 - d) * expC, p.exponent_relaxation_sigma - 1.0)
        Pmax = np.power(f["M_thick"], p.exponent_compression_factor) * p.compression_factor * expC
        with np.errstate(divide="ignore", invalid="ignore"):
            tildeP = np.where(sn < 0, np.minimum(1.0, -Pmax / sn), 0.0)
        mult = np.minimum(1.0 - 1e-12, tv / (tv + dt * (1.0 - tildeP)))
        el = p.young * (1.0 - d) * expC
        nu = p.nu0
        D = np.array([[1, nu, 0], [nu, 1, 0], [0, 0, (1 - nu) / 2]]) / (1 - nu * nu)
        eps = [e0, e1, e2]
        new = []
        for i in range(3):
            acc = s[i].copy()
            for j in range(3):
                acc = acc + dt * el * D[i, j] * eps[j]
            new.append(acc * mult)
        ss = np.hypot((new[0] - new[1]) / 2.0, new[2])
        sn = (new[0] + new[1]) * 0.5
        with np.errstate(divide="ignore", invalid="ignore"):
            dcrit = np.where(sn < -p.compr_strength, -p.compr_strength / sn, f["M_Cohesion"] / (ss + p.tan_phi * sn))
        hit = (0.0 < dcrit) & (dcrit < 1.0)
        rtd = np.sqrt(el) / (self.delta_x * np.sqrt(2.0 * (1.0 + nu) * RHOI))
        fac = np.where(hit, (1.0 - dcrit) * dt * rtd, 0.0)
        dnew = d + (1.0 - d) * fac
        new = [a - a * fac for a in new]
        dnew = np.maximum(0.0, dnew - dt / f["M_time_relaxation_damage"] * expC)
        f["M_damage"] = np.where(ice, dnew, 0.0)
        f["M_sigma"] = [np.where(ice, a, 0.0) for a in new]

    # FE.cpp:10649-10726
    def update_sigma_vp(self, ralpha1, ralpha2):
        f, p, t, nn = self.f, self.p, self.t, self.nn
        vt = f["M_VT"]
        u, v = vt[t], vt[t + nn]
        e11 = (self.dxN * u).sum(1)
        e22 = (self.dyN * v).sum(1)
        e12 = (0.5 * (self.dxN * v + self.dyN * u)).sum(1)
        re2 = 1.0 / (p.evp_e * p.evp_e)
        eps1, eps2 = e11 + e22, e11 - e22
        delta = np.sqrt(eps1 * eps1 + (eps2 * eps2 + 4 * e12 * e12) * re2)
        P = p.evp_Pstar * np.exp(-p.evp_C * (1.0 - f["M_conc"]))
        zeta = P / (delta + p.evp_dmin)
        s = f["M_sigma"]
        s1, s2 = s[0] + s[1], s[0] - s[1]
        s1 = s1 + ralpha1 * (zeta * (eps1 - delta) - s1)
        s2 = s2 + ralpha2 * (zeta * eps2 * re2 - s2)
        s12 = s[2] + ralpha2 * (zeta * e12 * re2 - s[2])
        ice = f["M_thick"] != 0.0
        f["M_sigma"] = [np.where(ice, 0.5 * (s1 + s2), 0.0), np.where(ice, 0.5 * (s1 - s2), 0.0), np.where(ice, s12, 0.0)]

    # one pass of the loop body FE.cpp:10420-10553
    def substep(self):
        f, p, t, nn = self.f, self.p, self.t, self.nn
        dte = p.dtime_step / p.substeps
        if p.dynamics_type == 0:
            self.update_sigma_damage(dte)
        elif p.dynamics_type == 3:
            T = p.dtime_step / 3.0
            self.update_sigma_vp(0.5 * dte / T, 0.5 * dte / T * p.evp_e * p.evp_e)
        else:
            self.update_sigma_vp(1.0 / p.mevp_alpha, 1.0 / p.mevp_alpha)
        s = f["M_sigma"]
        g = self.grad_ssh.copy()
        vol = f["M_thick"] * self.surface
        live = ~self.dir & (self.node_mass != 0.0)
        for i in range(3):
            ok = live[t[:, i]]
            np.subtract.at(g, t[ok, i], (vol * (s[0] * self.dxN[:, i] + s[2] * self.dyN[:, i]))[ok])
            np.subtract.at(g, t[ok, i] + nn, (vol * (s[2] * self.dxN[:, i] + s[1] * self.dyN[:, i]))[ok])
        vt = f["M_VT"]
        u, v = vt[:nn].copy(), vt[nn:].copy()
        oc = f["M_ocean"]
        if p.dynamics_type == 4:
            b = p.mevp_beta + 1.0
            delu, delv, dtep = (self.VTM[:nn] - u) / b, (self.VTM[nn:] - v) / b, dte / b
        else:
            delu = delv = np.zeros(nn)
            dtep = dte
        min_m = RHOI * p.min_h
        dom = dtep / np.maximum(min_m, self.node_mass)
        cp = RHOW * p.quad_drag_coef_water * np.hypot(oc[:nn] - u, oc[nn:] - v)
        cosa, sina = np.cos(p.ocean_turning_angle_rad), np.sin(p.ocean_turning_angle_rad)
        sgn = np.copysign(sina, self.lat)
        tau_b = self.cbu / (np.hypot(u, v) + p.basal_u0)
        alpha = 1.0 + dom * (cp * cosa + tau_b)
        beta = dtep * self.fcor + dom * cp * sgn
        rden = 1.0 / (alpha * alpha + beta * beta)
        tx = self.tau_a[:nn] + cp * (oc[:nn] * cosa - oc[nn:] * sgn)
        ty = self.tau_a[nn:] + cp * (oc[nn:] * cosa + oc[:nn] * sgn)
        gx, gy = g[:nn] * self.rlmass, g[nn:] * self.rlmass
        un = (alpha * u + beta * v + dom * (alpha * (gx + tx) + beta * (gy + ty)) + alpha * delu + beta * delv) * rden
        vn = (alpha * v - beta * u + dom * (alpha * (gy + ty) - beta * (gx + tx)) + alpha * delv - beta * delu) * rden
        vt[:nn] = np.where(live, un, u)
        vt[nn:] = np.where(live, vn, v)
        if p.dynamics_type != 4:
            self.move(dte)

    def move(self, dt):
        f, nn = self.f, self.nn
        um_p = f["M_UM"].copy()
        f["M_UM"] += dt * f["M_VT"]
        f["M_UT"] += dt * f["M_VT"]
        nm = np.concatenate([self.neumann, self.neumann + nn]).astype(np.int64)
        f["M_UM"][nm] = um_p[nm]

    def run(self, nsub):
        for _ in range(nsub):
            self.substep()
        if self.p.dynamics_type == 4:
            self.move(self.p.dtime_step)
        return self.f

    # FE.cpp:10578-10640: 50 Jacobi sweeps over the open-water nodes, ice-ocean stress, open-water mesh move
    def smooth_and_tauw(self, nodal_connectivity):
        f, p, nn = self.f, self.p, self.nn
        nc = np.asarray(nodal_connectivity)
        cnt = nc[:, -1].astype(np.int64)
        nbr = nc[:, :-1].astype(np.int64) - 1
        ow = np.nonzero(~self.dir & (self.node_mass == 0.0))[0]
        vt = f["M_VT"]
        valid = np.arange(nbr.shape[1])[None, :] < cnt[ow][:, None]
        idx = np.where(valid, nbr[ow], 0)
        for _ in range(50):
            old = vt.copy()
            vt[ow] = np.where(valid, old[idx], 0.0).sum(1) / cnt[ow]
            vt[ow + nn] = np.where(valid, old[idx + nn], 0.0).sum(1) / cnt[ow]
        oc = f["M_ocean"]
        ui, vi = 0.5 * (vt[:nn] + self.VTM[:nn]), 0.5 * (vt[nn:] + self.VTM[nn:])
        cp = RHOW * p.quad_drag_coef_water * np.hypot(oc[:nn] - ui, oc[nn:] - vi)
        self.tau_w = np.concatenate([cp * (ui - oc[:nn]), cp * (vi - oc[nn:])])
        um_p = f["M_UM"].copy()
        both = np.concatenate([ow, ow + nn])
        f["M_UM"][both] += p.dtime_step * vt[both]
        f["M_UT"][both] += p.dtime_step * vt[both]
        nm = np.concatenate([self.neumann, self.neumann + nn]).astype(np.int64)
        f["M_UM"][nm] = um_p[nm]

    # FE.cpp:3919-4132 (pure Lagrangian update; the sst/sss diffusion is thermodynamics, not on the path)
    def update(self, thick_myi, conc_myi, ridge_ratio):
        f, p, t, nn = self.f, self.p, self.t, self.nn
        um = f["M_UM"]
        vx, vy = self.x[t] + um[t], self.y[t] + um[t + nn]
        jac = (vx[:, 1] - vx[:, 0]) * (vy[:, 2] - vy[:, 0]) - (vx[:, 2] - vx[:, 0]) * (vy[:, 1] - vy[:, 0])
        new_surface = 0.5 * np.abs(jac)
        is_neu = np.zeros(nn, bool)
        is_neu[self.neumann] = True
        upd = (f["M_conc"] > 0.0) & ~is_neu[t].any(1)
        conc, thick, snow = f["M_conc"].copy(), f["M_thick"].copy(), f["M_snow_thick"].copy()
        hmyi, cmyi, rr = np.array(thick_myi, float), np.array(conc_myi, float), np.array(ridge_ratio, float)
        cy, hy, hsy = f["M_conc_young"].copy(), f["M_h_young"].copy(), f["M_hs_young"].copy()
        old_conc = conc.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.where(upd, self.surface / new_surface, 1.0)
            conc *= r; thick *= r; snow *= r; hmyi *= r
            sig = [s * r for s in f["M_sigma"]]
            rr = np.where(upd, 1.0 - (1.0 - rr) * np.minimum(1.0, conc) / (old_conc * r), rr)
            if self.young_ice:
                hy *= r; cy *= r; hsy *= r
            if p.equal_ridging:
                cmyi = np.where(upd, cmyi * (np.minimum(1.0, conc) / old_conc), cmyi)
            else:
                cmyi = np.where(upd, np.minimum(cmyi * r, 1.0), cmyi)
        ow = 1.0 - conc
        if self.young_ice:
            ow = ow - cy
        ow = np.clip(ow, 0.0, 1.0)
        ncy = np.zeros(self.ne); del_c = np.zeros(self.ne)
        if self.young_ice:
            has = cy > 0.0
            ncy = np.where(has, np.minimum(1.0, np.maximum(0.0, 1.0 - conc - ow)), 0.0)
            ridge = has & (conc > p.min_c) & (thick > p.min_h) & (ncy < cy)
            with np.errstate(divide="ignore", invalid="ignore"):
                nh, nhs = ncy * hy / cy, ncy * hsy / cy
            newice = np.where(ridge, hy - nh, 0.0)
            newsnow = np.where(ridge, hsy - nhs, 0.0)
            del_c = np.where(ridge, (cy - ncy) / 10.0, 0.0)
            with np.errstate(divide="ignore", invalid="ignore"):
                rr = np.where(ridge, 1.0 - (1.0 - rr) * thick / (thick + newice), rr)
            hy = np.where(ridge, nh, np.where(has, hy, 0.0))
            hsy = np.where(ridge, nhs, np.where(has, hsy, 0.0))
            thick = thick + newice
            snow = snow + newsnow
        conc = np.minimum(1.0, np.maximum(0.0, 1.0 - ncy - ow + del_c))
        if self.young_ice:
            ncy = np.maximum(0.0, np.minimum(ncy, 1.0 - conc))
            cy = ncy
        pos = conc > 0.0
        with np.errstate(divide="ignore", invalid="ignore"):
            hh = np.minimum(thick / conc, 50.0)
            conc = np.where(pos, np.minimum(1.0 - ncy, thick / hh), conc)
        rr = np.where(pos, rr, 0.0); thick = np.where(pos, thick, 0.0); snow = np.where(pos, snow, 0.0)
        conc = np.where(conc > 0, conc, 0.0); thick = np.where(thick > 0, thick, 0.0)
        hmyi = np.where(hmyi > 0, hmyi, 0.0); snow = np.where(snow > 0, snow, 0.0)
        cap = conc + cy if (p.newice_type == 4 and p.use_young_ice_in_myi_reset) else conc
        new_cmyi = np.maximum(0.0, np.minimum(cmyi, cap))
        return dict(M_conc=conc, M_thick=thick, M_snow_thick=snow, M_thick_myi=hmyi, M_conc_myi=new_cmyi,
                    M_ridge_ratio=rr, M_conc_young=cy, M_h_young=hy, M_hs_young=hsy, M_sigma=sig, M_surface=new_surface,
                    D_del_ci_ridge_myi=new_cmyi - cmyi)
