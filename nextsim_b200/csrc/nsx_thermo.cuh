// nsx_thermo.cuh -- FiniteElement::thermo(dt) as ONE element-wise function.
//
// The reference runs three loops over the elements -- OWBulkFluxes (FE.cpp:5032-5159), IABulkFluxes for old and for young
// ice (6148-6353), then the slab loop of thermo() (5278-6133) -- but every statement only touches element i (plus the
// three nodes of element i for the wind and ice-ocean speed), so they fuse into one pass: one thread per element, every
// field read once and written once.  Statement order, constants and libm calls follow the reference text line by line
// (citations in the comments); nothing is re-associated, so that the SAME function compiled for the host
// (tests/cpp/thermo_host.cpp) reproduces the reference's own compiled bodies (oracle/ref_fe) bit for bit, and the device
// instance differs from it only by the accuracy of CUDA's exp / pow / log / cbrt / atan (1-2 ulp).
//
// NSX_HD marks host+device functions; the header has no other CUDA dependency.
#pragma once
#include <cmath>

#include "../../include/nsx.h"

#if defined(__CUDACC__)
#define NSX_HD __host__ __device__ __forceinline__
#else
#define NSX_HD inline
#endif

namespace nsx {
namespace thermo {

// model/constants.hpp:12-86
namespace phys {
constexpr double C = 2100., cmin = 1e-12, cpa = 1000.5, cpv = 1860., cpw = 4186.84, eps = 0.996, g = 9.8, hmin = 0.01,
                 ki = 2.0334, Lf = 333.55e3, Lv0 = 2.5e6, Ra_dry = 287.058, Ra_vap = 461.5, rhoi = 917., rhow = 1025.,
                 rhos = 330., si = 5., sigma_sb = 5.67E-8, tfrwK = 273.15, vonKarman = 0.4, rhoa = 1.22, Gamma_d = 0.0098;
}
constexpr double days_in_sec = 86400.;

// options + the per-step scalars thermo() derives before its loops (FE.cpp:5180-5216, 6160-6205)
struct Params {
    NsxThermoParams o;
    int dt;                                 // thermo(int dt)
    double ddt;
    int step_in_day, num_steps_in_day;      // FE.cpp:5668-5672
    int midnight;                           // std::fmod(M_current_time, 1.) == 0.
    int is_0915, is_0801, is_reset_date;    // date_string_md == "0915" / "0801" / age.reset_date
    double timeT, timeS, rh0, rPhiF, qi, qs, h_young_max_sharp;
    // IABulkFluxes constants (FE.cpp:6171-6205)
    double z0, Linvrange, Bm, C1, C2, C3, Bm2, C4, C5, C6, C7, D1, D2, D3, D4, D5, lambda_u, lambda_h;
};

// every array thermo() touches, one pointer per reference member.  The X-lists give the member name without its prefix:
// forcing and state are M_<name>, diagnostics D_<name> (D_pond_fraction is state: IABulkFluxes reads it, meltPonds writes it).
#define NSX_THERMO_FORCING(X) \
    X(tair) X(mixrat) X(dair) X(sphuma) X(mslp) X(Qsw_in) X(Qlw_in) X(tcc) X(precip) X(snowfall) X(snowfr) X(mld) \
    X(ocean_temp) X(ocean_salt) X(conc_upd)
// the ice state the dynamics also owns (NsxFields on the C ABI)
#define NSX_THERMO_ICE(X) \
    X(conc) X(thick) X(snow_thick) X(conc_young) X(h_young) X(hs_young) X(ridge_ratio) X(conc_myi) X(thick_myi) X(drag_ui) \
    X(drag_ui_young) X(time_relaxation_damage)
#define NSX_THERMO_STATE(X) \
    X(sst) X(sss) X(tice0) X(tice1) X(tice2) X(tsurf_young) X(del_vi_tend) X(freeze_days) X(freeze_onset) X(conc_summer) \
    X(thick_summer) X(fyi_fraction) X(age_det) X(age) X(pond_volume) X(lid_volume) X(drag_ti) X(drag_ti_young)
#define NSX_THERMO_DIAG(X) \
    X(tau_ow) X(Qa) X(Qsw) X(Qlw) X(Qsh) X(Qlh) X(Qo) X(Qnosun) X(Qsw_ocean) X(Qassim) X(delS) X(fwflux_ice) X(fwflux) X(brine) \
    X(evap) X(rain) X(vice_melt) X(del_vi_young) X(del_hi) X(del_hi_young) X(newice) X(mlt_top) X(mlt_bot) X(snow2ice) X(albedo) \
    X(sialb) X(del_ci_mlt_myi) X(del_vi_mlt_myi) X(del_ci_rplnt_myi) X(del_vi_rplnt_myi)

struct Arrays {
    int ne, nn;
    const int *en0, *en1, *en2;             // 0-based node ids of the element (M_elements[i].indices[j]-1)
    const double *wind, *VT, *ocean;        // nodal, [u | v]
#define X(n) double* n;
    NSX_THERMO_FORCING(X)                   // read only
    NSX_THERMO_ICE(X)                       // in / out
    NSX_THERMO_STATE(X)                     // in / out
    double* pond_fraction;                  // in / out (D_pond_fraction)
    NSX_THERMO_DIAG(X)                      // out
#undef X
};

// reference member name -> slot of Arrays (nullptr for an unknown name); `shared` says the field is one of NsxFields'
inline double** field_slot(Arrays& A, const char* name, bool* shared = nullptr)
{
    auto eq = [](const char* a, const char* b) { while (*a && *a == *b) { ++a; ++b; } return *a == *b; };
    if (shared) *shared = false;
#define X(n) if (eq(name, "M_" #n)) return &A.n;
    NSX_THERMO_FORCING(X)
    NSX_THERMO_STATE(X)
#undef X
    if (eq(name, "D_pond_fraction")) return &A.pond_fraction;
#define X(n) if (eq(name, "D_" #n)) return &A.n;
    NSX_THERMO_DIAG(X)
#undef X
    if (shared) *shared = true;
#define X(n) if (eq(name, "M_" #n)) return &A.n;
    NSX_THERMO_ICE(X)
#undef X
    if (shared) *shared = false;
    return nullptr;
}

NSX_HD double dmax(double a, double b) { return (a < b) ? b : a; }      // std::max
NSX_HD double dmin(double a, double b) { return (b < a) ? b : a; }      // std::min

// FE.cpp:6432-6448
NSX_HD double freezingPoint(Params const& P, double sss)
{
    if (P.o.freezingpoint_type == 0) return -P.o.freezingpoint_mu * sss;
    return (-0.0575 + 1.710523e-3 * sqrt(sss) - 2.154996e-4 * sss) * sss;
}

// FE.cpp:6359-6370
NSX_HD double windSpeedElement(Arrays const& A, int i)
{
    int const n[3] = {A.en0[i], A.en1[i], A.en2[i]};
    double wspd = 0.;
    for (int j = 0; j < 3; ++j) wspd += hypot(A.wind[n[j]], A.wind[n[j] + A.nn]);
    return wspd / 3.;
}

// FE.cpp:6376-6389
NSX_HD double incomingLongwave(Params const& P, Arrays const& A, int i)
{
    if (P.o.have_Qlw_in) return A.Qlw_in[i];
    double taa = A.tair[i] + phys::tfrwK;
    return phys::sigma_sb * pow(taa, 4) * (1. - 0.261 * exp(-7.77e-4 * pow(taa - phys::tfrwK, 2))) * (1. + 0.275 * A.tcc[i]);
}

// FE.cpp:4966-5019.  scheme 0 atmosphere, 1 water, 2 ice; returns sphum, *dsphumdT (ice only)
NSX_HD double specificHumidity(Params const& P, Arrays const& A, int scheme, int i, double temp, double* dsphumdT)
{
    double Aa = 7.2e-4, B = 3.20e-6, Cc = 5.9e-10;
    double a = 6.1121e2, b = 18.729, c = 257.87, d = 227.3;
    double const alpha = 0.62197, beta = 0.37803;
    double salinity = 0.;
    *dsphumdT = 0.;
    if (scheme == 0) {
        if (P.o.have_sphuma) return dmax(0., A.sphuma[i]);
        if (P.o.have_mixrat) return A.mixrat[i] / (1. + A.mixrat[i]);
        temp = A.dair[i];
        salinity = 0;
    } else if (scheme == 1) {
        temp = A.sst[i];
        return 640380. / phys::rhoa * exp(-5107.4 / (temp + phys::tfrwK));
    } else {
        Aa = 2.2e-4, B = 3.83e-6, Cc = 6.4e-10;
        a = 6.1115e2, b = 23.036, c = 279.82, d = 333.7;
        salinity = 0;
    }
    double f = 1. + Aa + A.mslp[i] * 1e-2 * (B + Cc * temp * temp);
    double est = a * exp((b - temp / d) * temp / (temp + c)) * (1 - 5.37e-4 * salinity);
    double sphum = alpha * f * est / (A.mslp[i] - beta * f * est);
    if (scheme == 2) {
        double dfdT = 2. * Cc * B * temp;
        double destdT = (b * c * d - temp * (2. * c + temp)) / (d * pow(c + temp, 2)) * est;
        *dsphumdT = alpha * A.mslp[i] * (f * destdT + est * dfdT) / pow(A.mslp[i] - beta * est * f, 2);
    }
    return sphum;
}

// FE.cpp:6454-6535; returns albedo, *pen_sw.  alb_scheme outside 1..4 is rejected on the host.
NSX_HD double albedoFn(double Tsurf, double hs, double frac_pnd, int alb_scheme, double alb_ice, double alb_sn, double alb_pnd,
                       double I_0, double* pen_sw)
{
    double albedo;
    if (alb_scheme == 1 || alb_scheme == 2) {
        if (hs > 0.) {
            if (alb_scheme == 2) albedo = dmin(alb_sn, alb_ice + (alb_sn - alb_ice) * hs / 0.2);
            else albedo = alb_sn;
            *pen_sw = 0.;
        } else {
            albedo = alb_ice;
            *pen_sw = I_0;
        }
    } else if (alb_scheme == 3) {
        double albi, albs;
        if (Tsurf > -1.) {
            albi = alb_ice - 0.075 * (Tsurf + 1.);
            albs = alb_sn - 0.124 * (Tsurf + 1.);
        } else {
            albi = alb_ice;
            albs = alb_sn;
        }
        double frac_sn = hs / (hs + 0.02);
        albedo = frac_sn * albs + frac_pnd * alb_pnd + (1. - frac_sn - frac_pnd) * albi;
        *pen_sw = (1. - frac_sn - frac_pnd) * I_0;
    } else {
        double frac_sn = hs / (hs + 0.02);
        double albs;
        if (Tsurf > -1.) albs = alb_sn - 0.124 * (Tsurf + 1.);
        else albs = alb_sn;
        albedo = frac_sn * albs + frac_pnd * alb_pnd + (1. - frac_sn - frac_pnd) * alb_ice;
        *pen_sw = (1. - frac_sn - frac_pnd) * I_0;
    }
    return albedo;
}

struct IceFlux { double Qia, Qlw, Qsw, Qlh, Qsh, I, subl, dQiadT, alb_tot; };

// one element of IABulkFluxes (FE.cpp:6207-6352); drag_ui / drag_ti are updated in place like the reference's ModelVariables
NSX_HD IceFlux iaBulkFluxes(Params const& P, Arrays const& A, int i, double Tsurf, double snow_thick, double conc,
                            double& drag_ui, double& drag_ti, bool bulk_for_young)
{
    IceFlux F;
    double Qlw_out = phys::eps * phys::sigma_sb * pow(Tsurf + phys::tfrwK, 4);
    double dQlwdT = 4. * phys::eps * phys::sigma_sb * pow(Tsurf + phys::tfrwK, 3);

    double dsphumidT, dummy;
    double sphumi = specificHumidity(P, A, 2, i, Tsurf, &dsphumidT);
    double sphuma = specificHumidity(P, A, 0, i, 0., &dummy);

    double tairK = A.tair[i] + phys::tfrwK;
    double tsurfK = Tsurf + phys::tfrwK;
    double rhoair = A.mslp[i] / (phys::Ra_dry * tairK) * (1. - sphuma * (1. - phys::Ra_vap / phys::Ra_dry));
    double wspeed = windSpeedElement(A, i);
    const double Tpot = tairK + phys::Gamma_d * P.o.zref_temp;
    const double retv = 0.6078;
    const double ch = 3.;

    if (!P.o.force_neutral_atmosphere) {
        const double ustar = sqrt(drag_ui) * wspeed;
        const double Tvirt = Tpot * (1. + retv * sphuma);
        const double mixrat = sphuma / (1. - sphuma);
        const double wTpot = drag_ti * wspeed * (tsurfK - Tpot);
        const double wr = drag_ti * wspeed * (sphumi - sphuma) / ((1. - sphumi) * (1. - sphuma));
        const double wTvirt = wTpot * (1. + retv * mixrat) + retv * Tpot * wr;
        const double Linv = dmax(-P.Linvrange, dmin(P.Linvrange, -phys::vonKarman * phys::g * wTvirt / (ustar * ustar * ustar * Tvirt)));
        const double zetam = P.o.zref_wind * Linv;
        const double zetah = P.o.zref_temp * Linv;
        double psim, psih;
        if (Linv >= 0) {
            const double x = cbrt(1. + zetam);
            psim = P.C1 * (x - 1.) + P.C2 * (2. * log((x + P.Bm) * P.C3) - log((x * x - x * P.Bm + P.Bm2) * P.C4)
                                              + P.C5 * (atan((2. * x - P.Bm) * P.C6) - P.C7));
            psih = P.D1 * log(1. + ch * zetah + zetah * zetah) + P.D2 * (log((2. * zetah + P.D3) / (2. * zetah + P.D4)) - P.D5);
        } else {
            double x = sqrt(sqrt(1. - 16. * zetam));
            psim = 2. * log(0.5 * (1. + x)) + log(0.5 * (1. + x * x)) - 2. * atan(x) + 0.5 * 3.14159265358979323846;
            x = sqrt(sqrt(1. - 16. * zetah));
            psih = 2. * log(0.5 * (1. + x * x));
        }
        drag_ui = phys::vonKarman / (P.lambda_u - psim);
        drag_ui *= drag_ui;
        drag_ti = phys::vonKarman / (P.lambda_h - psih);
        drag_ti *= drag_ti;
    }

    F.Qsh = drag_ti * rhoair * phys::cpa * wspeed * (tsurfK - Tpot);
    double dQshdT = drag_ti * rhoair * phys::cpa * wspeed;
    double Lsub = phys::Lf + phys::Lv0 - 240. - 290. * Tsurf - 4. * Tsurf * Tsurf;
    F.Qlh = drag_ti * rhoair * Lsub * wspeed * (sphumi - sphuma);
    double dQlhdT = drag_ti * Lsub * rhoair * wspeed * dsphumidT;
    F.dQiadT = dQlwdT + dQshdT + dQlhdT;
    F.subl = dmax(0., F.Qlh / Lsub);

    double hs;
    if (conc > 0) hs = snow_thick / conc;
    else hs = 0;
    double pen_sw;
    double pond_fraction;
    if (A.pond_fraction[i] > 0. && A.lid_volume[i] / A.pond_fraction[i] <= 0.05) pond_fraction = A.pond_fraction[i];
    else pond_fraction = 0.;
    if (bulk_for_young) pond_fraction = 0.;
    F.alb_tot = albedoFn(Tsurf, hs, pond_fraction, P.o.alb_scheme, P.o.alb_ice, P.o.alb_sn, P.o.alb_ponds, P.o.I_0, &pen_sw);
    F.Qsw = -A.Qsw_in[i] * (1. - F.alb_tot) * (1. - pen_sw);
    F.I = A.Qsw_in[i] * (1. - F.alb_tot) * pen_sw;
    F.Qlw = Qlw_out - incomingLongwave(P, A, i);
    F.Qia = F.Qsw + F.Qlw + F.Qsh + F.Qlh;
    return F;
}

// FE.cpp:6396-6428
NSX_HD double iceOceanHeatflux(Params const& P, Arrays const& A, int cpt, double sst, double sss, double mld, double dt)
{
    double const Tbot = freezingPoint(P, sss);
    if (P.o.Qio_type == 0) return (sst - Tbot) * phys::rhow * phys::cpw * mld / dt;
    int const n[3] = {A.en0[cpt], A.en1[cpt], A.en2[cpt]};
    double welt_oce_ice = 0.;
    for (int i = 0; i < 3; ++i)
        welt_oce_ice += hypot(A.VT[n[i]] - A.ocean[n[i]], A.VT[n[i] + A.nn] - A.ocean[n[i] + A.nn]);
    double norm_Voce_ice = welt_oce_ice / 3.;
    return (sst - Tbot) * norm_Voce_ice * P.o.Csens_io * phys::rhow * phys::cpw;
}

// FE.cpp:6633-6853
NSX_HD void thermoWinton(Params const& P, double dt, double conc, double voli, double vols, double snowfall, double Qia,
                         double dQiadT, double I, double subl, double Tbot, double& Qio, double& hi, double& hs, double& hi_old,
                         double& del_hi, double& del_hs_mlt, double& mlt_hi_top, double& mlt_hi_bot, double& del_hi_s2i,
                         double& Tsurf, double& T1, double& T2)
{
    double const qi = phys::Lf * phys::rhoi;
    double const qs = phys::Lf * phys::rhos;
    double const Crho = phys::C * phys::rhoi;
    double const Tfr_ice = -P.o.freezingpoint_mu * phys::si;
    double const M_ks = P.o.ks;

    if (conc <= 0. || voli <= 0.) {
        hi = 0.; hs = 0.; hi_old = 0.; del_hi = 0.;
        Tsurf = Tfr_ice; T1 = Tfr_ice; T2 = Tfr_ice;
        return;
    }
    hi = voli / conc;
    hi_old = hi;
    hs = vols / conc;
    double const Tfr_surf = (hs > 0) ? 0. : Tfr_ice;

    double K12 = 4 * phys::ki * M_ks / (M_ks * hi + 4 * phys::ki * hs);
    double A = Qia - Tsurf * dQiadT;
    double B = dQiadT;
    double K32 = 2 * phys::ki / hi;

    double A1 = hi * Crho / (2 * dt) + K32 * (4 * dt * K32 + hi * Crho) / (6 * dt * K32 + hi * Crho) + K12 * B / (K12 + B);
    double B1 = -hi / (2 * dt) * (Crho * T1 + qi * Tfr_ice / T1) - I
                - K32 * (4 * dt * K32 * Tbot + hi * Crho * T2) / (6 * dt * K32 + hi * Crho) + A * K12 / (K12 + B);
    double C1 = hi * qi * Tfr_ice / (2 * dt);

    T1 = -(B1 + sqrt(B1 * B1 - 4 * A1 * C1)) / (2 * A1);
    Tsurf = (K12 * T1 - A) / (K12 + B);

    double Msurf = 0.;
    if (Tsurf > Tfr_surf) {
        Tsurf = Tfr_surf;
        A1 += K12 - K12 * B / (K12 + B);
        B1 -= K12 * Tsurf + A * K12 / (K12 + B);
        T1 = -(B1 + sqrt(B1 * B1 - 4 * A1 * C1)) / (2 * A1);
        Msurf = K12 * (T1 - Tsurf) - (A + B * Tsurf);
    }
    T2 = (2 * dt * K32 * (T1 + 2 * Tbot) + hi * Crho * T2) / (6 * dt * K32 + hi * Crho);

    double h1 = hi / 2.;
    double h2 = hi / 2.;
    double E1 = Crho * (T1 - Tfr_ice) - qi * (1 - Tfr_ice / T1);
    double E2 = Crho * (T2 - Tfr_ice) - qi;

    hs += snowfall / phys::rhos * dt;

    if (subl * dt <= hs * phys::rhos)
        hs -= subl * dt / phys::rhos;
    else if (subl * dt - hs * phys::rhos <= h1 * phys::rhoi) {
        h1 -= (subl * dt - hs * phys::rhos) / phys::rhoi;
        hs = 0.;
    } else if (subl * dt - h1 * phys::rhoi - hs * phys::rhos <= h2 * phys::rhoi) {
        h2 -= (subl * dt - h1 * phys::rhoi - hs * phys::rhos) / phys::rhoi;
        h1 = 0.;
        hs = 0.;
    } else {
        h2 = 0.; h1 = 0.; hs = 0.;          // "All the ice has sublimated" (the reference logs a warning)
    }
    mlt_hi_top = dmax(0., h1 + h2 - hi_old);

    double Mbot = Qio - 4 * phys::ki * (Tbot - T2) / hi;

    del_hs_mlt = 0;
    if (Mbot <= 0.) {
        double Ebot = Crho * (Tbot - Tfr_ice) - qi;
        double delh2 = Mbot * dt / Ebot;
        T2 = (delh2 * Tbot + h2 * T2) / (delh2 + h2);
        h2 += delh2;
    } else {
        double delh2 = -dmin(-Mbot * dt / E2, h2);
        double delh1 = -dmin(dmax(-(Mbot * dt + E2 * h2) / E1, 0.), h1);
        del_hs_mlt = -dmin(dmax((Mbot * dt + E2 * h2 + E1 * h1) / qs, 0.), hs);
        if (h2 + h1 + hs - delh2 - delh1 - del_hs_mlt <= 0.)
            Qio -= dmax(Mbot * dt - qs * hs + E1 * h1 + E2 * h2, 0.) / dt;
        hs += del_hs_mlt;
        h1 += delh1;
        h2 += delh2;
        mlt_hi_bot += delh1 + delh2;
    }

    del_hs_mlt -= dmin(Msurf * dt / qs, hs);
    double delh1 = -dmin(dmax(-(Msurf * dt - qs * hs) / E1, 0.), h1);
    double delh2 = -dmin(dmax(-(Msurf * dt - qs * hs + E1 * h1) / E2, 0.), h2);
    if (h2 + h1 + hs - delh2 - delh1 - del_hs_mlt <= 0.)
        Qio -= dmax(Msurf * dt - qs * hs + E1 * h1 + E2 * h2, 0.) / dt;

    hs += del_hs_mlt;
    h1 += delh1;
    h2 += delh2;
    mlt_hi_top += delh1 + delh2;

    double freeboard = (hi * (phys::rhow - phys::rhoi) - hs * phys::rhos) / phys::rhow;
    if (P.o.flooding && freeboard < 0) {
        hs += dmin(freeboard * phys::rhoi / phys::rhos, 0.);
        double delh1b = dmax(-freeboard, 0.);
        double f1 = 1 - delh1b / (delh1b + h1);
        double Tbar = f1 * (T1 + qi * Tfr_ice / (Crho * T1)) + (1 - f1) * Tfr_ice;
        T1 = (Tbar - sqrt(Tbar * Tbar - 4 * Tfr_ice * qi / Crho)) / 2.;
        h1 += delh1b;
        del_hi_s2i += delh1b;
    }
    hi = h1 + h2;

    if (h2 > h1) {
        double f1 = h1 / hi * 2.;
        double Tbar = f1 * (T1 + qi * Tfr_ice / (Crho * T1)) + (1 - f1) * T2;
        T1 = (Tbar - sqrt(Tbar * Tbar - 4 * Tfr_ice * qi / Crho)) / 2.;
    } else if (hi > 0.) {
        double f1 = (2. * h1 - hi) / hi;
        T2 = f1 * (T1 + qi * Tfr_ice / (Crho * T1)) + (1 - f1) * T2;
        if (T2 > Tfr_ice) {
            mlt_hi_top -= hi / 4 * Crho * (T2 - Tfr_ice) * T1 / (qi * T1 + (Crho * T1 - qi) * (Tfr_ice - T1));
            mlt_hi_bot -= hi / 4 * Crho * (T2 - Tfr_ice) * T1 / (qi * T1 + (Crho * T1 - qi) * (Tfr_ice - T1));
            hi -= hi / 2 * Crho * (T2 - Tfr_ice) * T1 / (qi * T1 + (Crho * T1 - qi) * (Tfr_ice - T1));
            T2 = Tfr_ice;
        }
    }
    del_hi = hi - hi_old;

    if (hi < phys::hmin) {
        Qio -= (-qs * hs + (E1 + E2) * hi / 2.) / dt;
        if (del_hi < 0.) {
            mlt_hi_top *= -hi_old / del_hi;
            mlt_hi_bot *= -hi_old / del_hi;
        }
        del_hi_s2i = 0.;
        del_hi = -hi_old;
        hi = 0.; hs = 0.;
        Tsurf = Tfr_ice; T1 = Tfr_ice; T2 = Tfr_ice;
    }
}

// FE.cpp:6860-6962
NSX_HD void thermoIce0(Params const& P, double dt, double conc, double voli, double vols, double snowfall, double Qia,
                       double dQiadT, double I, double subl, double Tbot, double& Qio, double& hi, double& hs, double& hi_old,
                       double& del_hi, double& del_hs_mlt, double& mlt_hi_top, double& mlt_hi_bot, double& del_hi_s2i, double& Tsurf)
{
    double const qi = phys::Lf * phys::rhoi;
    double const qs = phys::Lf * phys::rhos;
    double const Tfr_ice = -P.o.freezingpoint_mu * phys::si;
    double const beta = 0.4;
    double const gamma = 1.065;
    double const M_ks = P.o.ks;

    if (conc <= 0. || voli <= 0.) {
        hi = 0.; hi_old = 0.; hs = 0.;
        Tsurf = Tfr_ice;
        del_hi = 0.;
        return;
    }
    hi = voli / conc;
    hi_old = hi;
    hs = vols / conc;

    double Qic, del_hb, del_ht, draft;
    double const Qia_mod = Qia + (1. - beta) * I;

    Qic = M_ks * (Tbot - Tsurf) / (hs + M_ks * hi / phys::ki) * gamma;
    Tsurf = Tsurf + (Qic - Qia_mod) / (M_ks / (hs + M_ks * hi / phys::ki) + dQiadT);

    if (hs > 0.) Tsurf = dmin(0., Tsurf);
    else Tsurf = dmin(-P.o.freezingpoint_mu * phys::si, Tsurf);

    del_hs_mlt = dmin(Qia_mod - Qic, 0.) * dt / qs;
    hs += del_hs_mlt - subl * dt / phys::rhos;
    del_ht = dmin(hs, 0.) * qs / qi;
    hs = dmax(0., hs);
    hs += snowfall / phys::rhos * dt;

    del_hb = (Qic - Qio) * dt / qi;

    del_hi = del_ht + del_hb;
    hi = hi + del_hi;
    mlt_hi_top = dmin(del_ht, 0.);
    mlt_hi_bot = dmin(del_hb, 0.);

    draft = (hi * phys::rhoi + hs * phys::rhos) / phys::rhow;
    if (P.o.flooding && draft > hi) {
        del_hi_s2i += draft - hi;
        hs = hs - (draft - hi) * phys::rhoi / phys::rhos;
        hi = draft;
    }

    if (hi < phys::hmin) {
        if (del_hi < 0.) {
            mlt_hi_top *= -hi_old / del_hi;
            mlt_hi_bot *= -hi_old / del_hi;
        }
        del_hi_s2i = 0.;
        del_hi = -hi_old;
        Qio = Qio + hi * qi / dt + hs * qs / dt;
        hi = 0.; hs = 0.;
        Tsurf = Tfr_ice;
    }
}

// FE.cpp:6538-6627
NSX_HD void meltPonds(Params const& P, Arrays const& A, int cpt, double dt, double hi, double hs, double iceSurfaceMelt,
                      double snowMelt, double Qia, double rain, double roff, double dep2frac)
{
    const double hIceMin = 0.1;
    const double concMin = 0.1;
    const double max_lid_thickness = 0.3;
    const double min_lid_thickness = 1e-3;
    const double ice_to_water = phys::rhoi / phys::rhow;
    const double snow_to_water = phys::rhos / phys::rhow;
    const double water_to_ice = phys::rhow / phys::rhoi;

    double const availableWater = -iceSurfaceMelt * ice_to_water - snowMelt * snow_to_water + rain / phys::rhow * dt;
    A.pond_volume[cpt] += (1 - roff) * availableWater * A.conc[cpt];

    if (A.pond_volume[cpt] <= 0. || A.conc[cpt] <= concMin || A.thick[cpt] / A.conc[cpt] <= hIceMin) {
        A.pond_volume[cpt] = 0.;
        A.lid_volume[cpt] = 0.;
        A.pond_fraction[cpt] = 0.;
        return;
    }
    A.pond_fraction[cpt] = sqrt(A.pond_volume[cpt] / dep2frac);
    A.pond_fraction[cpt] = dmin(A.pond_fraction[cpt], 1. - hs / (hs + 0.2));
    double pond_depth = dmin(dep2frac * A.pond_fraction[cpt], 0.9 * hi);
    A.pond_volume[cpt] = pond_depth * A.pond_fraction[cpt];
    pond_depth = dmax(0.05, pond_depth);
    A.pond_fraction[cpt] = dmin(A.pond_fraction[cpt], (A.lid_volume[cpt] + A.pond_volume[cpt]) / pond_depth);

    double delLidVolume = 0;
    if (A.lid_volume[cpt] > 0. && A.pond_fraction[cpt] > 1e-11) {
        const double TPond = -P.o.freezingpoint_mu * phys::si;
        const double lidThickness = dmax(min_lid_thickness, dmin(max_lid_thickness, A.lid_volume[cpt] * water_to_ice / A.pond_fraction[cpt]));
        const double Qic = (TPond - A.tice0[cpt]) / lidThickness * phys::ki;
        const double delLidThickness = (dmin(Qia - Qic, 0.) + Qic) * dt / (phys::rhoi * phys::Lf);
        delLidVolume = delLidThickness * ice_to_water * A.pond_fraction[cpt];
        delLidVolume = dmax(delLidVolume, -A.lid_volume[cpt]);
    } else if (Qia > 0.) {
        delLidVolume = dt * Qia / (phys::rhoi * phys::Lf) * ice_to_water;
    }
    A.lid_volume[cpt] += delLidVolume;
    A.pond_volume[cpt] -= delLidVolume;
    if (A.pond_volume[cpt] <= 0. || A.lid_volume[cpt] * water_to_ice / A.pond_fraction[cpt] >= max_lid_thickness) {
        A.lid_volume[cpt] = 0.;
        A.pond_volume[cpt] = 0.;
        A.pond_fraction[cpt] = 0.;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// thermo() for element i
// ---------------------------------------------------------------------------------------------------------------------
NSX_HD void thermo_element(Params const& P, Arrays const& A, int i)
{
    NsxThermoParams const& o = P.o;
    double const ddt = P.ddt;
    int const dt = P.dt;
    double const qi = P.qi, qs = P.qs;
    bool const young = o.ice_cat_young != 0;
    double mld = o.constant_mld;

    // ---- OWBulkFluxes, element i (FE.cpp:5101-5158) ----
    double Qow, Qlw_ow, Qsw_ow, Qlh_ow, Qsh_ow, evap;
    {
        double dummy;
        double sphuma = specificHumidity(P, A, 0, i, 0., &dummy);
        double sphumw = specificHumidity(P, A, 1, i, 0., &dummy);
        double rhoair = A.mslp[i] / (phys::Ra_dry * (A.tair[i] + phys::tfrwK)) * (1. - sphuma * (1. - phys::Ra_vap / phys::Ra_dry));
        double wspeed = windSpeedElement(A, i);
        Qsh_ow = o.drag_ocean_t * rhoair * (phys::cpa + sphuma * phys::cpv) * wspeed * (A.sst[i] - A.tair[i]);
        double Lv = phys::Lv0 - 2.36418e3 * A.sst[i] + 1.58927 * A.sst[i] * A.sst[i] - 6.14342e-2 * pow(A.sst[i], 3.);
        Qlh_ow = dmax(o.drag_ocean_q * phys::rhoa * Lv * wspeed * (sphumw - sphuma), 0.);
        evap = Qlh_ow / Lv;
        double drag_ocean_m = 1e-3 * dmax(1., dmin(2., 0.61 + 0.063 * wspeed));
        A.tau_ow[i] = rhoair * drag_ocean_m;

        Qsw_ow = -A.Qsw_in[i] * (1. - o.ocean_albedo);
        double Qlw_out = phys::eps * phys::sigma_sb * pow(A.sst[i] + phys::tfrwK, 4.);
        Qlw_ow = Qlw_out - incomingLongwave(P, A, i);
        Qow = Qlw_ow + Qsh_ow + Qlh_ow;
        Qow += Qsw_ow;
    }

    // ---- IABulkFluxes over old ice and over young ice (FE.cpp:5248-5275) ----
    IceFlux Fi = iaBulkFluxes(P, A, i, A.tice0[i], A.snow_thick[i], A.conc[i], A.drag_ui[i], A.drag_ti[i], false);
    IceFlux Fy;
    Fy.Qia = Fy.Qlw = Fy.Qsw = Fy.Qlh = Fy.Qsh = Fy.I = Fy.subl = Fy.dQiadT = Fy.alb_tot = 0.;
    if (young)
        Fy = iaBulkFluxes(P, A, i, A.tsurf_young[i], A.hs_young[i], A.conc_young[i], A.drag_ui_young[i], A.drag_ti_young[i], true);

    // ---- the slab loop body (FE.cpp:5279-6132) ----
    double hi = 0., hi_old = 0., hs = 0.;
    double hi_young = 0., hi_young_old = 0., hs_young = 0.;
    double del_hi = 0., del_hi_young = 0.;
    double Qdw = 0., Fdw = 0.;
    double Qio = 0., Qio_young = 0.;
    double Qassm = 0.;

    double const old_vol = A.thick[i];
    double const old_snow_vol = A.snow_thick[i];
    (void)old_snow_vol;
    double const old_conc = A.conc[i];
    double old_h_young = 0., old_hs_young = 0., old_conc_young = 0.;
    if (young) {
        old_h_young = A.h_young[i];
        old_conc_young = A.conc_young[i];
        old_hs_young = A.hs_young[i];
    }
    (void)old_h_young; (void)old_hs_young;
    double const old_conc_tot = old_conc + old_conc_young;
    double const old_ow_fraction = 1. - old_conc_tot;

    double tmp_snowfall = 0.;
    if (o.have_snowfr) tmp_snowfall = A.precip[i] * A.snowfr[i];
    else if (o.have_snowfall) tmp_snowfall = A.snowfall[i];
    else if (A.tair[i] < 0) tmp_snowfall = A.precip[i];
    tmp_snowfall = dmax(0., tmp_snowfall);

    if (o.have_mld) mld = A.mld[i];

    if (o.ocean_constant) {
        Qdw = o.Qdw_const;
        Fdw = o.Fdw_const;
    } else {
        Qdw = -(A.sst[i] - A.ocean_temp[i]) * mld * phys::rhow * phys::cpw / P.timeT;
        double const delS = A.sss[i] - A.ocean_salt[i];
        Fdw = delS * mld * phys::rhow / (P.timeS * A.sss[i] - ddt * delS);
    }

    Qio = iceOceanHeatflux(P, A, i, A.sst[i], A.sss[i], mld, dt);
    if (young) Qio_young = Qio;
    const double tfrw = freezingPoint(P, A.sss[i]);

    double del_hs_mlt = 0, mlt_hi_top = 0, mlt_hi_bot = 0, del_hi_s2i = 0;
    if (o.thermo_type == 0)
        thermoIce0(P, ddt, A.conc[i], A.thick[i], A.snow_thick[i], tmp_snowfall, Fi.Qia, Fi.dQiadT, Fi.I, Fi.subl, tfrw,
                   Qio, hi, hs, hi_old, del_hi, del_hs_mlt, mlt_hi_top, mlt_hi_bot, del_hi_s2i, A.tice0[i]);
    else
        thermoWinton(P, ddt, A.conc[i], A.thick[i], A.snow_thick[i], tmp_snowfall, Fi.Qia, Fi.dQiadT, Fi.I, Fi.subl, tfrw,
                     Qio, hi, hs, hi_old, del_hi, del_hs_mlt, mlt_hi_top, mlt_hi_bot, del_hi_s2i, A.tice0[i], A.tice1[i], A.tice2[i]);

    double del_hs_young_mlt = 0, mlt_hi_top_young = 0, mlt_hi_bot_young = 0, del_hi_s2i_young = 0;
    if (young) {
        thermoIce0(P, ddt, A.conc_young[i], A.h_young[i], A.hs_young[i], tmp_snowfall, Fy.Qia, Fy.dQiadT, Fy.I, Fy.subl, tfrw,
                   Qio_young, hi_young, hs_young, hi_young_old, del_hi_young, del_hs_young_mlt, mlt_hi_top_young,
                   mlt_hi_bot_young, del_hi_s2i_young, A.tsurf_young[i]);
        A.h_young[i] = hi_young * old_conc_young;
        A.hs_young[i] = hs_young * old_conc_young;
    }

    double conc_pre_assim = old_conc + old_conc_young - A.conc_upd[i];
    if (o.use_assim_flux && (conc_pre_assim > 0) && (A.conc_upd[i] < 0))
        Qassm = (Qow * old_ow_fraction + Qio * old_conc + Qio_young * old_conc_young)
                * (pow(A.conc_upd[i] / conc_pre_assim + 1, o.assim_flux_exponent) - 1);

    double const tw_new = A.sst[i] - ddt * (Qow + Qassm) / (mld * phys::rhow * phys::cpw);

    double newice = 0;
    if (tw_new < tfrw) {
        newice = old_ow_fraction * (tfrw - tw_new) * mld * phys::rhow * phys::cpw / qi;
        Qow = -(tfrw - A.sst[i]) * mld * phys::rhow * phys::cpw / dt;
    }
    double const newice_stored = newice;

    double del_vi = newice + del_hi * old_conc;
    double mlt_vi_top = mlt_hi_top * old_conc;
    double mlt_vi_bot = mlt_hi_bot * old_conc;
    double del_vs_mlt = del_hs_mlt * old_conc;
    double snow2ice = del_hi_s2i * old_conc;
    double del_vi_young = 0.;
    if (young) {
        del_vi_young += del_hi_young * old_conc_young;
        del_vi += del_hi_young * old_conc_young;
        mlt_vi_top += mlt_hi_top_young * old_conc_young;
        mlt_vi_bot += mlt_hi_bot_young * old_conc_young;
        snow2ice += del_hi_s2i_young * old_conc_young;
        del_vs_mlt += del_hs_young_mlt * old_conc_young;
    }

    double del_c = 0.;
    double newsnow = 0.;

    switch (o.newice_type) {
        case 1:
            del_c = newice * P.rh0;
            break;
        case 2:
            if (hi_old > 0.) del_c = newice * o.PhiF / hi_old;
            else {
                if (newice > 0.) del_c = 1.;
                else del_c = 0.;
            }
            break;
        case 3: {
            double wspeed = windSpeedElement(A, i);
            double h0 = (1. + 0.1 * wspeed) / 15.;
            del_c = newice / dmax(P.rPhiF * hi_old, h0);
            break;
        }
        default:        // 4: young ice category (other values are rejected on the host)
            A.h_young[i] += newice;
            A.conc_young[i] = dmin(1. - A.conc[i], A.conc_young[i] + newice / o.h_young_min);
            newice = 0.;
            newsnow = 0.;
            if (A.conc_young[i] > 0.) {
                if (A.h_young[i] < o.h_young_min * A.conc_young[i]) {
                    A.conc_young[i] = A.h_young[i] / o.h_young_min;
                } else {
                    double const hiy = A.h_young[i] / A.conc_young[i];
                    if (hiy > P.h_young_max_sharp) {
                        double const hsy = dmax(0., A.hs_young[i] / A.conc_young[i]);
                        double tmp = A.conc_young[i] * (P.h_young_max_sharp - o.h_young_min) / (hiy - o.h_young_min);
                        del_c = dmax(0., A.conc_young[i] - tmp);
                        A.conc_young[i] = tmp;
                        tmp = A.conc_young[i] * P.h_young_max_sharp;
                        newice = dmax(0., A.h_young[i] - tmp);
                        A.h_young[i] = tmp;
                        tmp = A.conc_young[i] * hsy;
                        newsnow = dmax(0., A.hs_young[i] - tmp);
                        A.hs_young[i] = tmp;
                    }
                }
            } else {
                A.thick[i] += A.h_young[i];
                newice = A.h_young[i];
                newsnow = A.hs_young[i];
                A.h_young[i] = 0.;
                A.hs_young[i] = 0.;
            }
            break;
    }

    del_c = dmin(1. - A.conc[i], del_c);

    if (del_hi < 0.) {
        if (o.melt_type == 1) {
            if (A.conc[i] < 1.) del_c += del_hi * A.conc[i] * o.PhiM / hi_old;
            else del_c += 0.;
        } else {        // 2: Mellor and Kantha (89) (other values are rejected on the host)
            if (hi > 0.) {
                del_c += o.PhiM * (1. - A.conc[i]) * dmin(0., Qow) * ddt / (hi * qi + hs * qs);
                Qow *= (1. - o.PhiM);
            } else {
                del_c = -A.conc[i];
            }
        }
    }

    // ice age: freeze days (FE.cpp:5664-5699)
    bool use_young_ice_in_myi_reset = o.use_young_ice_in_myi_reset != 0;
    if (!o.reset_by_date) use_young_ice_in_myi_reset = false;
    if (P.step_in_day == 1) A.del_vi_tend[i] = 0.;
    A.del_vi_tend[i] += del_vi * ddt;
    if (P.step_in_day == P.num_steps_in_day) {
        if (A.del_vi_tend[i] > 0.) {
            A.freeze_days[i] += 1.;
        } else if (A.del_vi_tend[i] < 0.) {
            A.freeze_days[i] = 0.;
            double conc_summer = A.conc[i] + dmin(0., del_c);
            double thick_summer = A.thick[i] + dmin(0., del_vi);
            if (young && use_young_ice_in_myi_reset) {
                conc_summer += A.conc_young[i];
                thick_summer += A.h_young[i];
            }
            A.conc_summer[i] = dmax(0., dmin(1., conc_summer));
            A.thick_summer[i] = dmax(0., thick_summer);
        }
    }

    A.conc[i] += del_c;

    if (A.conc[i] >= phys::cmin) {
        hi = (hi * old_conc + newice) / A.conc[i];
        if (del_c < 0.) {
            Qow -= del_c * hs * qs / ddt;
        } else {
            hs = (hs * old_conc + newsnow) / A.conc[i];
        }
        if (o.thermo_type == 1) {
            double f1 = A.thick[i] / (A.thick[i] + newice);
            double Tbar = f1 * (A.tice1[i] - phys::Lf * o.freezingpoint_mu * phys::si / (phys::C * A.tice1[i])) + (1 - f1) * tfrw;
            A.tice1[i] = (Tbar - sqrt(Tbar * Tbar + 4 * o.freezingpoint_mu * phys::si * phys::Lf / phys::C)) / 2.;
            A.tice2[i] = f1 * A.tice2[i] + (1 - f1) * tfrw;
        }
    }

    if ((A.conc[i] < phys::cmin) || (hi < phys::hmin)) {
        Qow += A.conc[i] * hi * qi / ddt + A.conc[i] * hs * qs / ddt;
        A.conc[i] = 0.;
        A.tice0[i] = -o.freezingpoint_mu * phys::si;
        if (o.thermo_type == 1) {           // M_tice has three layers under Winton, one under the zero-layer scheme
            A.tice1[i] = -o.freezingpoint_mu * phys::si;
            A.tice2[i] = -o.freezingpoint_mu * phys::si;
        }
        hi = 0.;
        hs = 0.;
        A.ridge_ratio[i] = 0.;
    }

    A.thick[i] = hi * A.conc[i];
    A.snow_thick[i] = hs * A.conc[i];

    // ---- slab ocean (FE.cpp:5803-5846) ----
    double const rain_on_ice = dmax(0., A.precip[i] - tmp_snowfall);
    double rain = (1. - old_conc - old_conc_young) * A.precip[i] + (old_conc + old_conc_young) * rain_on_ice;
    double emp = evap * (1. - old_conc - old_conc_young) - rain;

    if (o.use_meltponds)
        meltPonds(P, A, i, ddt, hi, hs, mlt_hi_top, del_hs_mlt, Fi.Qia, rain_on_ice, o.meltpond_runoff_fraction, o.meltpond_depth_to_fraction);

    double Qio_mean = Qio * old_conc + Qio_young * old_conc_young;
    double Qow_mean = Qow * old_ow_fraction;

    A.sst[i] = A.sst[i] - ddt * (Qio_mean + Qow_mean - Qdw + Qassm) / (phys::rhow * phys::cpw * mld);

    double denominator = (mld * phys::rhow - del_vi * phys::rhoi - (del_vs_mlt * phys::rhos + (emp - Fdw) * ddt));
    denominator = (denominator > 1. * phys::rhow) ? denominator : 1. * phys::rhow;

    double const si_eff = dmin(A.sss[i], phys::si);
    double const delsss = ((A.sss[i] - si_eff) * phys::rhoi * del_vi + A.sss[i] * (del_vs_mlt * phys::rhos + (emp - Fdw) * ddt)) / denominator;
    A.sss[i] += delsss;

    if (A.thick[i] > old_vol) A.ridge_ratio[i] *= old_vol / A.thick[i];

    // ---- damage healing time (FE.cpp:5848-5882) ----
    if (o.temp_dep_healing) {
        if (A.thick[i] > 0.) {
            double deltaT;
            double Tbot = freezingPoint(P, A.sss[i]);
            double Cc;
            if (o.thermo_type == 0) {
                Cc = phys::ki * A.snow_thick[i] / (o.ks * A.thick[i]);
                deltaT = dmax(1e-36, Tbot - A.tice0[i]) / (1. + Cc);
            } else {
                Cc = phys::ki * A.snow_thick[i] / (o.ks * A.thick[i] / 4.);
                deltaT = dmax(1e-36, Tbot + Cc * (Tbot - A.tice1[i]) - A.tice0[i]) / (1. + Cc);
            }
            A.time_relaxation_damage[i] = dmax(o.time_relaxation_damage * o.deltaT_relaxation_damage / deltaT, ddt);
        } else {
            A.time_relaxation_damage[i] = 1e36;
        }
    }

    // ---- diagnostics (FE.cpp:5903-5984) ----
    A.Qa[i] = Fi.Qia * old_conc + Fy.Qia * old_conc_young + Qow * old_ow_fraction;
    A.Qsw[i] = Fi.Qsw * old_conc + Fy.Qsw * old_conc_young + Qsw_ow * old_ow_fraction;
    A.Qlw[i] = Fi.Qlw * old_conc + Fy.Qlw * old_conc_young + Qlw_ow * old_ow_fraction;
    A.Qsh[i] = Fi.Qsh * old_conc + Fy.Qsh * old_conc_young + Qsh_ow * old_ow_fraction;
    A.Qlh[i] = Fi.Qlh * old_conc + Fy.Qlh * old_conc_young + Qlh_ow * old_ow_fraction;
    A.Qo[i] = Qio_mean + Qow_mean;
    A.Qnosun[i] = Qio_mean + old_ow_fraction * (Qlw_ow + Qlh_ow + Qsh_ow);
    A.Qsw_ocean[i] = old_ow_fraction * Qsw_ow;
    A.Qassim[i] = Qassm;
    A.delS[i] = delsss * phys::rhow * mld * days_in_sec / o.dtime_step;
    A.fwflux_ice[i] = -1. / ddt * ((1. - 1e-3 * si_eff) * phys::rhoi * del_vi + phys::rhos * del_vs_mlt);
    A.fwflux[i] = A.fwflux_ice[i] - emp;
    A.brine[i] = -1e-3 * si_eff * phys::rhoi * del_vi / ddt;
    A.evap[i] = evap * (1. - old_conc - old_conc_young);
    A.rain[i] = rain;
    A.vice_melt[i] = del_vi * days_in_sec / ddt;
    A.del_vi_young[i] = del_vi_young * days_in_sec / ddt;
    A.del_hi[i] = del_hi * days_in_sec / ddt;
    A.del_hi_young[i] = del_hi_young * days_in_sec / ddt;
    A.newice[i] = newice_stored * days_in_sec / ddt;
    A.mlt_top[i] = mlt_vi_top * days_in_sec / ddt;
    A.mlt_bot[i] = mlt_vi_bot * days_in_sec / ddt;
    A.snow2ice[i] = snow2ice * days_in_sec / ddt;

    double sialb = old_conc * Fi.alb_tot;
    if (young) sialb += old_conc_young * Fy.alb_tot;
    A.albedo[i] = sialb + dmax(0., old_ow_fraction) * o.ocean_albedo;
    A.sialb[i] = (old_conc_tot > 0.) ? (sialb / old_conc_tot) : 0.;

    // ---- age / multi-year-ice tracers (FE.cpp:5986-6131) ----
    double del_vi_rplnt_myi = 0., del_ci_rplnt_myi = 0., del_vi_mlt_myi = 0., del_ci_mlt_myi = 0.;
    if (A.conc[i] < phys::cmin || A.thick[i] < A.conc[i] * phys::hmin) {
        A.fyi_fraction[i] = 0.;
        A.age_det[i] = 0.;
        A.age[i] = 0.;
        A.thick_myi[i] = 0.;
        A.conc_myi[i] = 0.;
        A.freeze_days[i] = 0.;
        A.freeze_onset[i] = 1.;
    } else {
        if (P.is_0915 && P.midnight) {
            A.fyi_fraction[i] = 0.;
        } else {
            double conc_fyi = A.fyi_fraction[i] + del_c;
            A.fyi_fraction[i] = dmax(0., dmin(1., conc_fyi));
        }
        double w_age = old_conc <= 0 ? 0. : dmin(old_conc / A.conc[i], 1.);
        A.age_det[i] = w_age * (A.age_det[i] + dt) + dmax((1 - w_age) * dt, 0.);
        w_age = old_vol <= 0 ? 0. : dmin(old_vol / A.thick[i], 1.);
        A.age[i] = w_age * (A.age[i] + dt) + dmax((1 - w_age) * dt, 0.);

        bool reset_myi = false;
        if (o.reset_by_date) {
            if (P.is_reset_date && P.midnight) reset_myi = true;
        } else {
            if (A.freeze_days[i] >= o.freeze_days_threshold) {
                if (A.freeze_onset[i] <= 0.5) {
                    reset_myi = true;
                    A.freeze_onset[i] = 1.;
                }
            }
        }
        if (P.is_0801 && P.midnight) {
            A.freeze_onset[i] = 0.;
            double ctot = A.conc[i];
            if (young) ctot += A.conc_young[i];
            if (ctot == 0.) A.freeze_onset[i] = 1.;
            double conc_summer = A.conc[i];
            double thick_summer = A.thick[i];
            if (young && use_young_ice_in_myi_reset) {
                conc_summer += A.conc_young[i];
                thick_summer += A.h_young[i];
            }
            A.conc_summer[i] = dmax(0., dmin(1., conc_summer));
            A.thick_summer[i] = dmax(0., thick_summer);
        }
        A.freeze_onset[i] = round(A.freeze_onset[i]);

        double old_conc_myi = A.conc_myi[i];
        double old_thick_myi = A.thick_myi[i];
        double c_myi_max = A.conc[i];
        double v_myi_max = A.thick[i];
        if (young && use_young_ice_in_myi_reset) {
            c_myi_max += A.conc_young[i];
            v_myi_max += A.h_young[i];
        }
        if (reset_myi) {
            if (!o.reset_by_date) {
                double c_myi_reset = dmax(A.conc_summer[i], A.conc_myi[i]);
                double v_myi_reset = dmax(A.thick_summer[i], A.thick_myi[i]);
                A.conc_myi[i] = dmin(c_myi_max, c_myi_reset);
                A.thick_myi[i] = dmin(v_myi_max, v_myi_reset);
            } else {
                A.conc_myi[i] = c_myi_max;
                A.thick_myi[i] = v_myi_max;
            }
            A.conc_myi[i] = dmax(0., dmin(1., A.conc_myi[i]));
            A.thick_myi[i] = dmax(0., A.thick_myi[i]);
            del_ci_rplnt_myi = A.conc_myi[i] - old_conc_myi;
            del_vi_rplnt_myi = A.thick_myi[i] - old_thick_myi;
        } else {
            if ((A.thick[i] < old_vol) && (old_conc > 0) && (old_vol > 0)) {
                if (o.equal_melting) {
                    double const del_c_ratio = dmin(A.conc[i] / old_conc, 1.);
                    double const del_v_ratio = dmin(A.thick[i] / old_vol, 1.);
                    del_ci_mlt_myi = dmin(0., A.conc_myi[i] * (del_c_ratio - 1.));
                    del_vi_mlt_myi = dmin(0., A.thick_myi[i] * (del_v_ratio - 1.));
                }
                A.conc_myi[i] = dmax(0., dmin(c_myi_max, A.conc_myi[i] + del_ci_mlt_myi));
                A.thick_myi[i] = dmax(0., dmin(v_myi_max, A.thick_myi[i] + del_vi_mlt_myi));
                del_ci_mlt_myi = A.conc_myi[i] - old_conc_myi;
                del_vi_mlt_myi = A.thick_myi[i] - old_thick_myi;
            }
        }
    }
    A.del_ci_mlt_myi[i] = del_ci_mlt_myi * days_in_sec / ddt;
    A.del_vi_mlt_myi[i] = del_vi_mlt_myi * days_in_sec / ddt;
    A.del_ci_rplnt_myi[i] = del_ci_rplnt_myi * days_in_sec / ddt;
    A.del_vi_rplnt_myi[i] = del_vi_rplnt_myi * days_in_sec / ddt;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: options -> Params (the scalars thermo() and IABulkFluxes derive before their loops)
// ---------------------------------------------------------------------------------------------------------------------
// nextsim time (decimal days since 1900-01-01 00:00, core/include/date.hpp) -> month, day
inline void month_day(double datenum, int& month, int& day)
{
    long z = (long)std::floor(datenum) + 693901L - 60L;
    long const era = (z >= 0 ? z : z - 146096) / 146097;
    unsigned long const doe = (unsigned long)(z - era * 146097);
    unsigned long const yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    unsigned long const doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    unsigned long const mp = (5 * doy + 2) / 153;
    day = (int)(doy - (153 * mp + 2) / 5 + 1);
    month = (int)(mp < 10 ? mp + 3 : mp - 9);
}

// model/options.cpp:272-449, 543-548 ([thermo], [ideal_simul], [age], [dynamics]) and FE.cpp:1186-1295
inline void params_defaults(NsxThermoParams& p)
{
    p = NsxThermoParams();
    p.thermo_type = 1; p.ocean_constant = 1; p.Qio_type = 0; p.freezingpoint_type = 0;
    p.newice_type = 4; p.melt_type = 2; p.alb_scheme = 3; p.flooding = 1;
    p.use_assim_flux = 0; p.temp_dep_healing = 0; p.use_meltponds = 0; p.force_neutral_atmosphere = 0;
    p.reset_by_date = 0; p.equal_melting = 1; p.use_young_ice_in_myi_reset = 1; p.ice_cat_young = 1;
    p.have_sphuma = p.have_mixrat = p.have_Qlw_in = p.have_snowfr = p.have_snowfall = p.have_mld = 0;
    p.reset_month = 9; p.reset_day = 15;
    p.dtime_step = 200.;
    p.ocean_nudge_timeT_days = 30.; p.ocean_nudge_timeS_days = 30.;
    p.Qdw_const = 0.; p.Fdw_const = 0.;
    p.hnull = 0.25; p.PhiF = 4.; p.PhiM = 0.5;
    p.assim_flux_exponent = 1.;
    p.constant_mld = 9.;
    p.I_0 = 0.30;
    p.freeze_days_threshold = 3.;
    p.meltpond_runoff_fraction = 0.2; p.meltpond_depth_to_fraction = 0.8;
    p.drag_ocean_t = 0.83e-3; p.drag_ocean_q = 1.5e-3;
    p.alb_ice = 0.538; p.alb_sn = 0.8256; p.alb_ponds = 0.30;
    p.zref_wind = 10.; p.zref_temp = 2.; p.limiting_lengthscale = 1.;
    p.quad_drag_coef_air = 0.0049;
    p.ocean_albedo = 0.07;
    p.ks = 0.3096;
    p.freezingpoint_mu = 0.055;
    p.Csens_io = 1e-3;
    p.time_relaxation_damage = 25. * 86400.;
    p.deltaT_relaxation_damage = 20.;
    p.h_young_min = 0.05; p.h_young_max = 0.5;
}

// the option values the reference rejects with std::logic_error (FE.cpp:5562-5565, 5653-5656, 6527-6529) or that need
// code this build does not have (OASIS: melt_type 3); nullptr when the options are usable
inline const char* validate(NsxThermoParams const& o, int dt)
{
    if (dt <= 0) return "thermo: dt must be positive";
    if (o.thermo_type != 0 && o.thermo_type != 1) return "thermo: setup.thermo-type must be 0 (zero-layer) or 1 (winton)";
    if (o.newice_type < 1 || o.newice_type > 4) return "Wrong newice_type";
    if (o.melt_type == 3) return "thermo: melt_type 3 needs the OASIS floe-size distribution, not built";
    if (o.melt_type < 1 || o.melt_type > 2) return "Wrong melt_type";
    if (o.alb_scheme < 1 || o.alb_scheme > 4) return "Wrong albedo_scheme";
    if ((o.newice_type == 4) != (o.ice_cat_young != 0)) return "thermo: ice_cat_young must be set exactly when newice_type == 4 (FE.cpp:1211-1214)";
    if (!(o.dtime_step > 0.)) return "thermo: simul.timestep must be positive";
    return nullptr;
}

inline Params make_params(NsxThermoParams const& o, int dt, double current_time)
{
    Params P;
    P.o = o;
    P.dt = dt;
    P.ddt = double(dt);                                                              // FE.cpp:5175
    P.timeT = days_in_sec * o.ocean_nudge_timeT_days;                                // :5179-5180
    P.timeS = days_in_sec * o.ocean_nudge_timeS_days;
    P.rh0 = 1. / o.hnull;                                                            // :5184-5185
    P.rPhiF = 1. / o.PhiF;
    P.qi = phys::Lf * phys::rhoi;                                                    // :5187-5188
    P.qs = phys::Lf * phys::rhos;
    P.h_young_max_sharp = .5 * (o.h_young_min + o.h_young_max);                      // :1198
    P.num_steps_in_day = (int)std::round(days_in_sec / o.dtime_step);                // :5668-5670
    P.step_in_day = 1 + (int)std::round(P.num_steps_in_day * std::fmod(current_time, 1.));
    P.midnight = std::fmod(current_time, 1.) == 0.;
    int m, d;
    month_day(current_time, m, d);                                                   // datenumToString(M_current_time, "%m%d"), :5216
    P.is_0915 = (m == 9 && d == 15);
    P.is_0801 = (m == 8 && d == 1);
    P.is_reset_date = (m == o.reset_month && d == o.reset_day);
    // IABulkFluxes, FE.cpp:6171-6205
    P.z0 = o.zref_wind * std::exp(-phys::vonKarman / std::sqrt(o.quad_drag_coef_air));
    P.Linvrange = 1. / o.limiting_lengthscale;
    double const am = 5.;
    double const bm = am / 6.5;
    P.Bm = std::cbrt((1 - bm) / bm);
    double const ah = 5., bh = 5., ch = 3.;
    double const Bh = std::sqrt(5);
    P.C1 = -3. * am / bm;
    P.C2 = 0.5 * am * P.Bm / bm;
    P.C3 = 1. / (1. + P.Bm);
    P.Bm2 = P.Bm * P.Bm;
    P.C4 = 1. / (1. - P.Bm + P.Bm2);
    double const sqrt3 = std::sqrt(3.);
    P.C5 = 2. * sqrt3;
    P.C6 = 1. / (sqrt3 * P.Bm);
    P.C7 = std::atan((2. - P.Bm) * P.C6);
    P.D1 = -0.5 * bh;
    P.D2 = -ah / Bh + 0.5 * bh * ch / Bh;
    P.D3 = ch - Bh;
    P.D4 = ch + Bh;
    P.D5 = std::log(P.D3 / P.D4);
    P.lambda_u = std::log(o.zref_wind / P.z0);
    P.lambda_h = std::log(o.zref_wind / P.z0);
    return P;
}

}  // namespace thermo
}  // namespace nsx
