// nsx_thermo.cuh -- FiniteElement::thermo(dt) as ONE element-wise function.
//
// The reference runs three loops over the elements -- OWBulkFluxes (FE.cpp:5032-5159), IABulkFluxes for old and for young
// ice (6148-6353), then the slab loop of thermo() (5278-6133) -- but every statement only touches element i (plus the
// three nodes of element i for the wind and ice-ocean speed), so they fuse into one pass: one thread per element, every
// field read once and written once.  Statement order and constants follow the reference text line by line (citations in
// the comments) and nothing is re-associated, so that the HOST build of this function (oracle/thermo_oracle.cpp) reproduces
// the reference's own compiled bodies (oracle/ref_fe) bit for bit (tests/test_thermo_cpu.py).
//
// The DEVICE build deviates from that text in four marked places, each worth <= 1-2 ulp per operation: CUDA's libm for
// exp / log / cbrt / atan; pow(x, 2|3|4) spelled as products and hypot as sqrt(u*u + v*v) (pw2..pw4, hyp); division by a
// divisor that is the same for all elements as multiplication by its host-computed reciprocal (UDiv); one reciprocal per
// element for divisors that divide several quantities (recip).  The kernel is built -fmad=false (nsx_thermo.cu).
// Quantities all three bulk-flux loops recompute from the same inputs (atmospheric humidity, air density, wind speed,
// incoming long-wave) are computed once -- same values, no rounding change.
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

// A divisor that is the same for every element (a physical constant, the time step): the host build divides, as the
// reference does; the device multiplies by the reciprocal computed once on the host (an IEEE double division is ~25
// instructions, a third of the ~150 divisions of one element have such a divisor).  <= 1 ulp apart, like the libm calls.
// -DNSX_THERMO_EXACT_DIV keeps the division on the device too.
struct UDiv { double c, rc; };
NSX_HD double operator/(double x, UDiv const& d)
{
#if defined(__CUDA_ARCH__) && !defined(NSX_THERMO_EXACT_DIV)
    return x * d.rc;
#else
    return x / d.c;
#endif
}

// The same for a divisor that belongs to the element but divides several quantities (the concentration, the denominators of
// the Winton temperature solve): one reciprocal on the device, plain divisions on the host.
NSX_HD UDiv recip(double c)
{
#if defined(__CUDA_ARCH__) && !defined(NSX_THERMO_EXACT_DIV)
    return UDiv{c, 1. / c};
#else
    return UDiv{c, 0.};
#endif
}

// options + the per-step scalars thermo() derives before its loops (FE.cpp:5180-5216, 6160-6205)
struct Params {
    NsxThermoParams o;
    int dt;                                 // thermo(int dt)
    double ddt;
    int step_in_day, num_steps_in_day;      // FE.cpp:5668-5672
    int midnight;                           // std::fmod(M_current_time, 1.) == 0.
    int is_0915, is_0801, is_reset_date;    // date_string_md == "0915" / "0801" / age.reset_date
    double timeT, timeS, rh0, rPhiF, qi, qs, h_young_max_sharp;
    UDiv u_ddt, u_2ddt, u_rhos, u_rhow, u_rhoi, u_qi, u_qs, u_ki, u_h_young_min, u_Crho;
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

#if defined(__CUDA_ARCH__) && defined(NSX_THERMO_FAST_MINMAX)
NSX_HD double dmax(double a, double b) { return fmax(a, b); }           // one DMNMX; differs from std::max for NaN and +-0 only
NSX_HD double dmin(double a, double b) { return fmin(a, b); }
#else
NSX_HD double dmax(double a, double b) { return (a < b) ? b : a; }      // std::max
NSX_HD double dmin(double a, double b) { return (b < a) ? b : a; }      // std::min
#endif
// std::pow(x, 2|3|4) and std::hypot of the reference.  The host build keeps the libm calls (bit for bit with the reference's
// build); the device spells them as products: CUDA's pow(double, double) is ~350 instructions, sixteen of them per element
// made the kernel instruction-bound (profiles/r2_thermo_v1.txt), and x*x differs from a correctly rounded pow by <= 1 ulp.
#if defined(__CUDA_ARCH__)
NSX_HD double pw2(double x) { return x * x; }
NSX_HD double pw3(double x) { return x * x * x; }
NSX_HD double pw4(double x) { double const y = x * x; return y * y; }
NSX_HD double hyp(double u, double v) { return sqrt(u * u + v * v); }
#else
NSX_HD double pw2(double x) { return std::pow(x, 2); }
NSX_HD double pw3(double x) { return std::pow(x, 3); }
NSX_HD double pw4(double x) { return std::pow(x, 4); }
NSX_HD double hyp(double u, double v) { return std::hypot(u, v); }
#endif

// exp / log / atan / cbrt: one out-of-line copy each on the device.  Inlined at their 9 call sites (x2 for the two
// IABulkFluxes instances) they made the kernel body larger than the SM's instruction cache (stall "no instruction" was the
// top stall reason, profiles/r2_thermo_v3.txt); arguments and results travel in registers, so the call costs a few cycles.
#if defined(__CUDA_ARCH__)
#define NSX_LIBM_WRAP(name) __device__ __noinline__ static double m_##name(double x) { return ::name(x); }
#else
#define NSX_LIBM_WRAP(name) inline double m_##name(double x) { return std::name(x); }
#endif
NSX_LIBM_WRAP(exp)
NSX_LIBM_WRAP(log)
NSX_LIBM_WRAP(atan)
NSX_LIBM_WRAP(cbrt)
#undef NSX_LIBM_WRAP

// one element's fields, held in registers between the loads at the top of thermo_element() and the stores at its end
struct Elem {
#define X(n) double n;
    NSX_THERMO_FORCING(X)
    NSX_THERMO_ICE(X)
    NSX_THERMO_STATE(X)
    double pond_fraction;
#undef X
};

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
    for (int j = 0; j < 3; ++j) wspd += hyp(A.wind[n[j]], A.wind[n[j] + A.nn]);
    return wspd / 3.;
}

// FE.cpp:6376-6389
NSX_HD double incomingLongwave(Params const& P, Elem const& E)
{
    if (P.o.have_Qlw_in) return E.Qlw_in;
    double taa = E.tair + phys::tfrwK;
    return phys::sigma_sb * pw4(taa) * (1. - 0.261 * m_exp(-7.77e-4 * pw2(taa - phys::tfrwK))) * (1. + 0.275 * E.tcc);
}

// FE.cpp:4966-5019.  scheme 0 atmosphere, 1 water, 2 ice; returns sphum, *dsphumdT (ice only)
NSX_HD double specificHumidity(Params const& P, Elem const& E, int scheme, double temp, double* dsphumdT)
{
    double Aa = 7.2e-4, B = 3.20e-6, Cc = 5.9e-10;
    double a = 6.1121e2, b = 18.729, c = 257.87, d = 227.3;
    double const alpha = 0.62197, beta = 0.37803;
    double salinity = 0.;
    *dsphumdT = 0.;
    if (scheme == 0) {
        if (P.o.have_sphuma) return dmax(0., E.sphuma);
        if (P.o.have_mixrat) return E.mixrat / (1. + E.mixrat);
        temp = E.dair;
        salinity = 0;
    } else if (scheme == 1) {
        temp = E.sst;
        return 640380. / phys::rhoa * m_exp(-5107.4 / (temp + phys::tfrwK));
    } else {
        Aa = 2.2e-4, B = 3.83e-6, Cc = 6.4e-10;
        a = 6.1115e2, b = 23.036, c = 279.82, d = 333.7;
        salinity = 0;
    }
    double f = 1. + Aa + E.mslp * 1e-2 * (B + Cc * temp * temp);
    double est = a * m_exp((b - temp / d) * temp / (temp + c)) * (1 - 5.37e-4 * salinity);
    double sphum = alpha * f * est / (E.mslp - beta * f * est);
    if (scheme == 2) {
        double dfdT = 2. * Cc * B * temp;
        double destdT = (b * c * d - temp * (2. * c + temp)) / (d * pw2(c + temp)) * est;
        *dsphumdT = alpha * E.mslp * (f * destdT + est * dfdT) / pw2(E.mslp - beta * est * f);
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
// what OWBulkFluxes and both IABulkFluxes calls each recompute from the same inputs (pure functions of element i):
// specificHumidity(ATMOSPHERE), the air density, windSpeedElement(), incomingLongwave()
struct Air { double sphuma, rhoair, wspeed, Qlw_in; };

// one element of IABulkFluxes (FE.cpp:6207-6352); drag_ui / drag_ti are updated in place like the reference's ModelVariables
NSX_HD IceFlux iaBulkFluxes(Params const& P, Elem const& E, Air const& air, double Tsurf, double snow_thick, double conc,
                            double& drag_ui, double& drag_ti, bool bulk_for_young)
{
    IceFlux F;
    double Qlw_out = phys::eps * phys::sigma_sb * pw4(Tsurf + phys::tfrwK);
    double dQlwdT = 4. * phys::eps * phys::sigma_sb * pw3(Tsurf + phys::tfrwK);

    double dsphumidT;
    double sphumi = specificHumidity(P, E, 2, Tsurf, &dsphumidT);
    double const sphuma = air.sphuma;                       // FE.cpp:6224-6225

    double tairK = E.tair + phys::tfrwK;
    double tsurfK = Tsurf + phys::tfrwK;
    double const rhoair = air.rhoair;                       // FE.cpp:6234: same expression as FE.cpp:5115
    double const wspeed = air.wspeed;                       // FE.cpp:6237
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
            const double x = m_cbrt(1. + zetam);
            psim = P.C1 * (x - 1.) + P.C2 * (2. * m_log((x + P.Bm) * P.C3) - m_log((x * x - x * P.Bm + P.Bm2) * P.C4)
                                              + P.C5 * (m_atan((2. * x - P.Bm) * P.C6) - P.C7));
            psih = P.D1 * m_log(1. + ch * zetah + zetah * zetah) + P.D2 * (m_log((2. * zetah + P.D3) / (2. * zetah + P.D4)) - P.D5);
        } else {
            double x = sqrt(sqrt(1. - 16. * zetam));
            psim = 2. * m_log(0.5 * (1. + x)) + m_log(0.5 * (1. + x * x)) - 2. * m_atan(x) + 0.5 * 3.14159265358979323846;
            x = sqrt(sqrt(1. - 16. * zetah));
            psih = 2. * m_log(0.5 * (1. + x * x));
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
    if (E.pond_fraction > 0. && E.lid_volume / E.pond_fraction <= 0.05) pond_fraction = E.pond_fraction;
    else pond_fraction = 0.;
    if (bulk_for_young) pond_fraction = 0.;
    F.alb_tot = albedoFn(Tsurf, hs, pond_fraction, P.o.alb_scheme, P.o.alb_ice, P.o.alb_sn, P.o.alb_ponds, P.o.I_0, &pen_sw);
    F.Qsw = -E.Qsw_in * (1. - F.alb_tot) * (1. - pen_sw);
    F.I = E.Qsw_in * (1. - F.alb_tot) * pen_sw;
    F.Qlw = Qlw_out - air.Qlw_in;
    F.Qia = F.Qsw + F.Qlw + F.Qsh + F.Qlh;
    return F;
}

// FE.cpp:6396-6428
NSX_HD double iceOceanHeatflux(Params const& P, Arrays const& A, int cpt, double sst, double sss, double mld, double /*dt: P.u_ddt*/)
{
    double const Tbot = freezingPoint(P, sss);
    if (P.o.Qio_type == 0) return (sst - Tbot) * phys::rhow * phys::cpw * mld / P.u_ddt;
    int const n[3] = {A.en0[cpt], A.en1[cpt], A.en2[cpt]};
    double welt_oce_ice = 0.;
    for (int i = 0; i < 3; ++i)
        welt_oce_ice += hyp(A.VT[n[i]] - A.ocean[n[i]], A.VT[n[i] + A.nn] - A.ocean[n[i] + A.nn]);
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
    UDiv const u_conc = recip(conc);
    hi = voli / u_conc;
    hi_old = hi;
    hs = vols / u_conc;
    double const Tfr_surf = (hs > 0) ? 0. : Tfr_ice;

    double K12 = 4 * phys::ki * M_ks / (M_ks * hi + 4 * phys::ki * hs);
    double A = Qia - Tsurf * dQiadT;
    double B = dQiadT;
    UDiv const u_hi = recip(hi);
    double K32 = 2 * phys::ki / u_hi;

    UDiv const u_D6 = recip(6 * dt * K32 + hi * Crho), u_KB = recip(K12 + B);
    double A1 = hi * Crho / P.u_2ddt + K32 * (4 * dt * K32 + hi * Crho) / u_D6 + K12 * B / u_KB;
    double B1 = -hi / P.u_2ddt * (Crho * T1 + qi * Tfr_ice / T1) - I
                - K32 * (4 * dt * K32 * Tbot + hi * Crho * T2) / u_D6 + A * K12 / u_KB;
    double C1 = hi * qi * Tfr_ice / P.u_2ddt;

    T1 = -(B1 + sqrt(B1 * B1 - 4 * A1 * C1)) / (2 * A1);
    Tsurf = (K12 * T1 - A) / u_KB;

    double Msurf = 0.;
    if (Tsurf > Tfr_surf) {
        Tsurf = Tfr_surf;
        A1 += K12 - K12 * B / u_KB;
        B1 -= K12 * Tsurf + A * K12 / u_KB;
        T1 = -(B1 + sqrt(B1 * B1 - 4 * A1 * C1)) / (2 * A1);
        Msurf = K12 * (T1 - Tsurf) - (A + B * Tsurf);
    }
    T2 = (2 * dt * K32 * (T1 + 2 * Tbot) + hi * Crho * T2) / u_D6;

    double h1 = hi / 2.;
    double h2 = hi / 2.;
    double E1 = Crho * (T1 - Tfr_ice) - qi * (1 - Tfr_ice / T1);
    double E2 = Crho * (T2 - Tfr_ice) - qi;

    hs += snowfall / P.u_rhos * dt;

    if (subl * dt <= hs * phys::rhos)
        hs -= subl * dt / P.u_rhos;
    else if (subl * dt - hs * phys::rhos <= h1 * phys::rhoi) {
        h1 -= (subl * dt - hs * phys::rhos) / P.u_rhoi;
        hs = 0.;
    } else if (subl * dt - h1 * phys::rhoi - hs * phys::rhos <= h2 * phys::rhoi) {
        h2 -= (subl * dt - h1 * phys::rhoi - hs * phys::rhos) / P.u_rhoi;
        h1 = 0.;
        hs = 0.;
    } else {
        h2 = 0.; h1 = 0.; hs = 0.;          // "All the ice has sublimated" (the reference logs a warning)
    }
    mlt_hi_top = dmax(0., h1 + h2 - hi_old);

    double Mbot = Qio - 4 * phys::ki * (Tbot - T2) / u_hi;

    del_hs_mlt = 0;
    if (Mbot <= 0.) {
        double Ebot = Crho * (Tbot - Tfr_ice) - qi;
        double delh2 = Mbot * dt / Ebot;
        T2 = (delh2 * Tbot + h2 * T2) / (delh2 + h2);
        h2 += delh2;
    } else {
        double delh2 = -dmin(-Mbot * dt / E2, h2);
        double delh1 = -dmin(dmax(-(Mbot * dt + E2 * h2) / E1, 0.), h1);
        del_hs_mlt = -dmin(dmax((Mbot * dt + E2 * h2 + E1 * h1) / P.u_qs, 0.), hs);
        if (h2 + h1 + hs - delh2 - delh1 - del_hs_mlt <= 0.)
            Qio -= dmax(Mbot * dt - qs * hs + E1 * h1 + E2 * h2, 0.) / P.u_ddt;
        hs += del_hs_mlt;
        h1 += delh1;
        h2 += delh2;
        mlt_hi_bot += delh1 + delh2;
    }

    del_hs_mlt -= dmin(Msurf * dt / P.u_qs, hs);
    double delh1 = -dmin(dmax(-(Msurf * dt - qs * hs) / E1, 0.), h1);
    double delh2 = -dmin(dmax(-(Msurf * dt - qs * hs + E1 * h1) / E2, 0.), h2);
    if (h2 + h1 + hs - delh2 - delh1 - del_hs_mlt <= 0.)
        Qio -= dmax(Msurf * dt - qs * hs + E1 * h1 + E2 * h2, 0.) / P.u_ddt;

    hs += del_hs_mlt;
    h1 += delh1;
    h2 += delh2;
    mlt_hi_top += delh1 + delh2;

    double freeboard = (hi * (phys::rhow - phys::rhoi) - hs * phys::rhos) / P.u_rhow;
    if (P.o.flooding && freeboard < 0) {
        hs += dmin(freeboard * phys::rhoi / P.u_rhos, 0.);
        double delh1b = dmax(-freeboard, 0.);
        double f1 = 1 - delh1b / (delh1b + h1);
        double Tbar = f1 * (T1 + qi * Tfr_ice / (Crho * T1)) + (1 - f1) * Tfr_ice;
        T1 = (Tbar - sqrt(Tbar * Tbar - 4 * Tfr_ice * qi / P.u_Crho)) / 2.;
        h1 += delh1b;
        del_hi_s2i += delh1b;
    }
    hi = h1 + h2;

    if (h2 > h1) {
        double f1 = h1 / hi * 2.;
        double Tbar = f1 * (T1 + qi * Tfr_ice / (Crho * T1)) + (1 - f1) * T2;
        T1 = (Tbar - sqrt(Tbar * Tbar - 4 * Tfr_ice * qi / P.u_Crho)) / 2.;
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
        Qio -= (-qs * hs + (E1 + E2) * hi / 2.) / P.u_ddt;
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
    UDiv const u_conc = recip(conc);
    hi = voli / u_conc;
    hi_old = hi;
    hs = vols / u_conc;

    double Qic, del_hb, del_ht, draft;
    double const Qia_mod = Qia + (1. - beta) * I;

    UDiv const u_res = recip(hs + M_ks * hi / P.u_ki);
    Qic = M_ks * (Tbot - Tsurf) / u_res * gamma;
    Tsurf = Tsurf + (Qic - Qia_mod) / (M_ks / u_res + dQiadT);

    if (hs > 0.) Tsurf = dmin(0., Tsurf);
    else Tsurf = dmin(-P.o.freezingpoint_mu * phys::si, Tsurf);

    del_hs_mlt = dmin(Qia_mod - Qic, 0.) * dt / P.u_qs;
    hs += del_hs_mlt - subl * dt / P.u_rhos;
    del_ht = dmin(hs, 0.) * qs / P.u_qi;
    hs = dmax(0., hs);
    hs += snowfall / P.u_rhos * dt;

    del_hb = (Qic - Qio) * dt / P.u_qi;

    del_hi = del_ht + del_hb;
    hi = hi + del_hi;
    mlt_hi_top = dmin(del_ht, 0.);
    mlt_hi_bot = dmin(del_hb, 0.);

    draft = (hi * phys::rhoi + hs * phys::rhos) / P.u_rhow;
    if (P.o.flooding && draft > hi) {
        del_hi_s2i += draft - hi;
        hs = hs - (draft - hi) * phys::rhoi / P.u_rhos;
        hi = draft;
    }

    if (hi < phys::hmin) {
        if (del_hi < 0.) {
            mlt_hi_top *= -hi_old / del_hi;
            mlt_hi_bot *= -hi_old / del_hi;
        }
        del_hi_s2i = 0.;
        del_hi = -hi_old;
        Qio = Qio + hi * qi / P.u_ddt + hs * qs / P.u_ddt;
        hi = 0.; hs = 0.;
        Tsurf = Tfr_ice;
    }
}

// FE.cpp:6538-6627
NSX_HD void meltPonds(Params const& P, Elem& E, double dt, double hi, double hs, double iceSurfaceMelt,
                      double snowMelt, double Qia, double rain, double roff, double dep2frac)
{
    const double hIceMin = 0.1;
    const double concMin = 0.1;
    const double max_lid_thickness = 0.3;
    const double min_lid_thickness = 1e-3;
    const double ice_to_water = phys::rhoi / P.u_rhow;
    const double snow_to_water = phys::rhos / P.u_rhow;
    const double water_to_ice = phys::rhow / P.u_rhoi;

    double const availableWater = -iceSurfaceMelt * ice_to_water - snowMelt * snow_to_water + rain / P.u_rhow * dt;
    E.pond_volume += (1 - roff) * availableWater * E.conc;

    if (E.pond_volume <= 0. || E.conc <= concMin || E.thick / E.conc <= hIceMin) {
        E.pond_volume = 0.;
        E.lid_volume = 0.;
        E.pond_fraction = 0.;
        return;
    }
    E.pond_fraction = sqrt(E.pond_volume / dep2frac);
    E.pond_fraction = dmin(E.pond_fraction, 1. - hs / (hs + 0.2));
    double pond_depth = dmin(dep2frac * E.pond_fraction, 0.9 * hi);
    E.pond_volume = pond_depth * E.pond_fraction;
    pond_depth = dmax(0.05, pond_depth);
    E.pond_fraction = dmin(E.pond_fraction, (E.lid_volume + E.pond_volume) / pond_depth);

    double delLidVolume = 0;
    if (E.lid_volume > 0. && E.pond_fraction > 1e-11) {
        const double TPond = -P.o.freezingpoint_mu * phys::si;
        const double lidThickness = dmax(min_lid_thickness, dmin(max_lid_thickness, E.lid_volume * water_to_ice / E.pond_fraction));
        const double Qic = (TPond - E.tice0) / lidThickness * phys::ki;
        const double delLidThickness = (dmin(Qia - Qic, 0.) + Qic) * dt / (phys::rhoi * phys::Lf);
        delLidVolume = delLidThickness * ice_to_water * E.pond_fraction;
        delLidVolume = dmax(delLidVolume, -E.lid_volume);
    } else if (Qia > 0.) {
        delLidVolume = dt * Qia / (phys::rhoi * phys::Lf) * ice_to_water;
    }
    E.lid_volume += delLidVolume;
    E.pond_volume -= delLidVolume;
    if (E.pond_volume <= 0. || E.lid_volume * water_to_ice / E.pond_fraction >= max_lid_thickness) {
        E.lid_volume = 0.;
        E.pond_volume = 0.;
        E.pond_fraction = 0.;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// thermo() for element i
// ---------------------------------------------------------------------------------------------------------------------
NSX_HD void thermo_core(Params const& P, Arrays const& A, int i, Elem& E)
{
    NsxThermoParams const& o = P.o;
    double const ddt = P.ddt;
    int const dt = P.dt;
    double const qi = P.qi, qs = P.qs;
    bool const young = o.ice_cat_young != 0;
    double mld = o.constant_mld;

    // ---- OWBulkFluxes, element i (FE.cpp:5101-5158) ----
    double Qow, Qlw_ow, Qsw_ow, Qlh_ow, Qsh_ow, evap;
    Air air;
    {
        double dummy;
        double sphuma = specificHumidity(P, E, 0, 0., &dummy);
        double sphumw = specificHumidity(P, E, 1, 0., &dummy);
        double rhoair = E.mslp / (phys::Ra_dry * (E.tair + phys::tfrwK)) * (1. - sphuma * (1. - phys::Ra_vap / phys::Ra_dry));
        double wspeed = windSpeedElement(A, i);
        air.sphuma = sphuma; air.rhoair = rhoair; air.wspeed = wspeed; air.Qlw_in = incomingLongwave(P, E);
        Qsh_ow = o.drag_ocean_t * rhoair * (phys::cpa + sphuma * phys::cpv) * wspeed * (E.sst - E.tair);
        double Lv = phys::Lv0 - 2.36418e3 * E.sst + 1.58927 * E.sst * E.sst - 6.14342e-2 * pw3(E.sst);
        Qlh_ow = dmax(o.drag_ocean_q * phys::rhoa * Lv * wspeed * (sphumw - sphuma), 0.);
        evap = Qlh_ow / Lv;
        double drag_ocean_m = 1e-3 * dmax(1., dmin(2., 0.61 + 0.063 * wspeed));
        A.tau_ow[i] = rhoair * drag_ocean_m;

        Qsw_ow = -E.Qsw_in * (1. - o.ocean_albedo);
        double Qlw_out = phys::eps * phys::sigma_sb * pw4(E.sst + phys::tfrwK);
        Qlw_ow = Qlw_out - air.Qlw_in;
        Qow = Qlw_ow + Qsh_ow + Qlh_ow;
        Qow += Qsw_ow;
    }

    // ---- IABulkFluxes over old ice and over young ice (FE.cpp:5248-5275) ----
    IceFlux Fi = iaBulkFluxes(P, E, air, E.tice0, E.snow_thick, E.conc, E.drag_ui, E.drag_ti, false);
    IceFlux Fy;
    Fy.Qia = Fy.Qlw = Fy.Qsw = Fy.Qlh = Fy.Qsh = Fy.I = Fy.subl = Fy.dQiadT = Fy.alb_tot = 0.;
    if (young)
        Fy = iaBulkFluxes(P, E, air, E.tsurf_young, E.hs_young, E.conc_young, E.drag_ui_young, E.drag_ti_young, true);

    // ---- the slab loop body (FE.cpp:5279-6132) ----
    double hi = 0., hi_old = 0., hs = 0.;
    double hi_young = 0., hi_young_old = 0., hs_young = 0.;
    double del_hi = 0., del_hi_young = 0.;
    double Qdw = 0., Fdw = 0.;
    double Qio = 0., Qio_young = 0.;
    double Qassm = 0.;

    double const old_vol = E.thick;                         // (old_snow_vol, old_h_young, old_hs_young of FE.cpp:5307-5317 are never read)
    double const old_conc = E.conc;
    double const old_conc_young = young ? E.conc_young : 0.;
    double const old_conc_tot = old_conc + old_conc_young;
    double const old_ow_fraction = 1. - old_conc_tot;

    // diagnostics that only need the fluxes and the old concentrations are written now (FE.cpp:5906-5924, 5972-5977),
    // which ends the live range of eight flux components
    A.Qsw[i] = Fi.Qsw * old_conc + Fy.Qsw * old_conc_young + Qsw_ow * old_ow_fraction;
    A.Qlw[i] = Fi.Qlw * old_conc + Fy.Qlw * old_conc_young + Qlw_ow * old_ow_fraction;
    A.Qsh[i] = Fi.Qsh * old_conc + Fy.Qsh * old_conc_young + Qsh_ow * old_ow_fraction;
    A.Qlh[i] = Fi.Qlh * old_conc + Fy.Qlh * old_conc_young + Qlh_ow * old_ow_fraction;
    double const Qnosun_ow = old_ow_fraction * (Qlw_ow + Qlh_ow + Qsh_ow);
    A.Qsw_ocean[i] = old_ow_fraction * Qsw_ow;
    {
        double sialb = old_conc * Fi.alb_tot;
        if (young) sialb += old_conc_young * Fy.alb_tot;
        A.albedo[i] = sialb + dmax(0., old_ow_fraction) * o.ocean_albedo;
        A.sialb[i] = (old_conc_tot > 0.) ? (sialb / old_conc_tot) : 0.;
    }

    double tmp_snowfall = 0.;
    if (o.have_snowfr) tmp_snowfall = E.precip * E.snowfr;
    else if (o.have_snowfall) tmp_snowfall = E.snowfall;
    else if (E.tair < 0) tmp_snowfall = E.precip;
    tmp_snowfall = dmax(0., tmp_snowfall);

    if (o.have_mld) mld = E.mld;

    if (o.ocean_constant) {
        Qdw = o.Qdw_const;
        Fdw = o.Fdw_const;
    } else {
        Qdw = -(E.sst - E.ocean_temp) * mld * phys::rhow * phys::cpw / P.timeT;
        double const delS = E.sss - E.ocean_salt;
        Fdw = delS * mld * phys::rhow / (P.timeS * E.sss - ddt * delS);
    }

    Qio = iceOceanHeatflux(P, A, i, E.sst, E.sss, mld, dt);
    if (young) Qio_young = Qio;
    const double tfrw = freezingPoint(P, E.sss);

    double del_hs_mlt = 0, mlt_hi_top = 0, mlt_hi_bot = 0, del_hi_s2i = 0;
    if (o.thermo_type == 0)
        thermoIce0(P, ddt, E.conc, E.thick, E.snow_thick, tmp_snowfall, Fi.Qia, Fi.dQiadT, Fi.I, Fi.subl, tfrw,
                   Qio, hi, hs, hi_old, del_hi, del_hs_mlt, mlt_hi_top, mlt_hi_bot, del_hi_s2i, E.tice0);
    else
        thermoWinton(P, ddt, E.conc, E.thick, E.snow_thick, tmp_snowfall, Fi.Qia, Fi.dQiadT, Fi.I, Fi.subl, tfrw,
                     Qio, hi, hs, hi_old, del_hi, del_hs_mlt, mlt_hi_top, mlt_hi_bot, del_hi_s2i, E.tice0, E.tice1, E.tice2);

    double del_hs_young_mlt = 0, mlt_hi_top_young = 0, mlt_hi_bot_young = 0, del_hi_s2i_young = 0;
    if (young) {
        thermoIce0(P, ddt, E.conc_young, E.h_young, E.hs_young, tmp_snowfall, Fy.Qia, Fy.dQiadT, Fy.I, Fy.subl, tfrw,
                   Qio_young, hi_young, hs_young, hi_young_old, del_hi_young, del_hs_young_mlt, mlt_hi_top_young,
                   mlt_hi_bot_young, del_hi_s2i_young, E.tsurf_young);
        E.h_young = hi_young * old_conc_young;
        E.hs_young = hs_young * old_conc_young;
    }

    double conc_pre_assim = old_conc + old_conc_young - E.conc_upd;
    if (o.use_assim_flux && (conc_pre_assim > 0) && (E.conc_upd < 0))
        Qassm = (Qow * old_ow_fraction + Qio * old_conc + Qio_young * old_conc_young)
                * (pow(E.conc_upd / conc_pre_assim + 1, o.assim_flux_exponent) - 1);

    double const tw_new = E.sst - ddt * (Qow + Qassm) / (mld * phys::rhow * phys::cpw);

    double newice = 0;
    if (tw_new < tfrw) {
        newice = old_ow_fraction * (tfrw - tw_new) * mld * phys::rhow * phys::cpw / P.u_qi;
        Qow = -(tfrw - E.sst) * mld * phys::rhow * phys::cpw / P.u_ddt;
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
            double wspeed = air.wspeed;                 // windSpeedElement(i), FE.cpp:5511
            double h0 = (1. + 0.1 * wspeed) / 15.;
            del_c = newice / dmax(P.rPhiF * hi_old, h0);
            break;
        }
        default:        // 4: young ice category (other values are rejected on the host)
            E.h_young += newice;
            E.conc_young = dmin(1. - E.conc, E.conc_young + newice / P.u_h_young_min);
            newice = 0.;
            newsnow = 0.;
            if (E.conc_young > 0.) {
                if (E.h_young < o.h_young_min * E.conc_young) {
                    E.conc_young = E.h_young / P.u_h_young_min;
                } else {
                    double const hiy = E.h_young / E.conc_young;
                    if (hiy > P.h_young_max_sharp) {
                        double const hsy = dmax(0., E.hs_young / E.conc_young);
                        double tmp = E.conc_young * (P.h_young_max_sharp - o.h_young_min) / (hiy - o.h_young_min);
                        del_c = dmax(0., E.conc_young - tmp);
                        E.conc_young = tmp;
                        tmp = E.conc_young * P.h_young_max_sharp;
                        newice = dmax(0., E.h_young - tmp);
                        E.h_young = tmp;
                        tmp = E.conc_young * hsy;
                        newsnow = dmax(0., E.hs_young - tmp);
                        E.hs_young = tmp;
                    }
                }
            } else {
                E.thick += E.h_young;
                newice = E.h_young;
                newsnow = E.hs_young;
                E.h_young = 0.;
                E.hs_young = 0.;
            }
            break;
    }

    del_c = dmin(1. - E.conc, del_c);

    if (del_hi < 0.) {
        if (o.melt_type == 1) {
            if (E.conc < 1.) del_c += del_hi * E.conc * o.PhiM / hi_old;
            else del_c += 0.;
        } else {        // 2: Mellor and Kantha (89) (other values are rejected on the host)
            if (hi > 0.) {
                del_c += o.PhiM * (1. - E.conc) * dmin(0., Qow) * ddt / (hi * qi + hs * qs);
                Qow *= (1. - o.PhiM);
            } else {
                del_c = -E.conc;
            }
        }
    }

    // tracer state is loaded where it is first needed: it would otherwise sit in registers through the flux and slab code
    E.del_vi_tend = A.del_vi_tend[i]; E.freeze_days = A.freeze_days[i]; E.ridge_ratio = A.ridge_ratio[i];
    // M_conc_summer / M_thick_summer: read by the freeze-days reset (FE.cpp:6083-6084), written on the last step of a day and
    // at the 1 August midnight (FE.cpp:5680-5697, 6059-6077); on other steps of a reset-by-date run they are not touched
    bool const summer_written = P.step_in_day == P.num_steps_in_day || (P.is_0801 && P.midnight);
    if (!o.reset_by_date || summer_written) { E.conc_summer = A.conc_summer[i]; E.thick_summer = A.thick_summer[i]; }
    // ice age: freeze days (FE.cpp:5664-5699)
    bool use_young_ice_in_myi_reset = o.use_young_ice_in_myi_reset != 0;
    if (!o.reset_by_date) use_young_ice_in_myi_reset = false;
    if (P.step_in_day == 1) E.del_vi_tend = 0.;
    E.del_vi_tend += del_vi * ddt;
    if (P.step_in_day == P.num_steps_in_day) {
        if (E.del_vi_tend > 0.) {
            E.freeze_days += 1.;
        } else if (E.del_vi_tend < 0.) {
            E.freeze_days = 0.;
            double conc_summer = E.conc + dmin(0., del_c);
            double thick_summer = E.thick + dmin(0., del_vi);
            if (young && use_young_ice_in_myi_reset) {
                conc_summer += E.conc_young;
                thick_summer += E.h_young;
            }
            E.conc_summer = dmax(0., dmin(1., conc_summer));
            E.thick_summer = dmax(0., thick_summer);
        }
    }

    E.conc += del_c;

    if (E.conc >= phys::cmin) {
        UDiv const u_c = recip(E.conc);
        hi = (hi * old_conc + newice) / u_c;
        if (del_c < 0.) {
            Qow -= del_c * hs * qs / P.u_ddt;
        } else {
            hs = (hs * old_conc + newsnow) / u_c;
        }
        if (o.thermo_type == 1) {
            double f1 = E.thick / (E.thick + newice);
            double Tbar = f1 * (E.tice1 - phys::Lf * o.freezingpoint_mu * phys::si / (phys::C * E.tice1)) + (1 - f1) * tfrw;
            E.tice1 = (Tbar - sqrt(Tbar * Tbar + 4 * o.freezingpoint_mu * phys::si * phys::Lf / phys::C)) / 2.;
            E.tice2 = f1 * E.tice2 + (1 - f1) * tfrw;
        }
    }

    if ((E.conc < phys::cmin) || (hi < phys::hmin)) {
        Qow += E.conc * hi * qi / P.u_ddt + E.conc * hs * qs / P.u_ddt;
        E.conc = 0.;
        E.tice0 = -o.freezingpoint_mu * phys::si;
        if (o.thermo_type == 1) {           // M_tice has three layers under Winton, one under the zero-layer scheme
            E.tice1 = -o.freezingpoint_mu * phys::si;
            E.tice2 = -o.freezingpoint_mu * phys::si;
        }
        hi = 0.;
        hs = 0.;
        E.ridge_ratio = 0.;
    }

    E.thick = hi * E.conc;
    E.snow_thick = hs * E.conc;

    // ---- slab ocean (FE.cpp:5803-5846) ----
    double const rain_on_ice = dmax(0., E.precip - tmp_snowfall);
    double rain = (1. - old_conc - old_conc_young) * E.precip + (old_conc + old_conc_young) * rain_on_ice;
    double emp = evap * (1. - old_conc - old_conc_young) - rain;

    if (o.use_meltponds)
        meltPonds(P, E, ddt, hi, hs, mlt_hi_top, del_hs_mlt, Fi.Qia, rain_on_ice, o.meltpond_runoff_fraction, o.meltpond_depth_to_fraction);

    double Qio_mean = Qio * old_conc + Qio_young * old_conc_young;
    double Qow_mean = Qow * old_ow_fraction;

    E.sst = E.sst - ddt * (Qio_mean + Qow_mean - Qdw + Qassm) / (phys::rhow * phys::cpw * mld);

    double denominator = (mld * phys::rhow - del_vi * phys::rhoi - (del_vs_mlt * phys::rhos + (emp - Fdw) * ddt));
    denominator = (denominator > 1. * phys::rhow) ? denominator : 1. * phys::rhow;

    double const si_eff = dmin(E.sss, phys::si);
    double const delsss = ((E.sss - si_eff) * phys::rhoi * del_vi + E.sss * (del_vs_mlt * phys::rhos + (emp - Fdw) * ddt)) / denominator;
    E.sss += delsss;

    if (E.thick > old_vol) E.ridge_ratio *= old_vol / E.thick;

    // ---- damage healing time (FE.cpp:5848-5882) ----
    if (o.temp_dep_healing) {
        if (E.thick > 0.) {
            double deltaT;
            double Tbot = freezingPoint(P, E.sss);
            double Cc;
            if (o.thermo_type == 0) {
                Cc = phys::ki * E.snow_thick / (o.ks * E.thick);
                deltaT = dmax(1e-36, Tbot - E.tice0) / (1. + Cc);
            } else {
                Cc = phys::ki * E.snow_thick / (o.ks * E.thick / 4.);
                deltaT = dmax(1e-36, Tbot + Cc * (Tbot - E.tice1) - E.tice0) / (1. + Cc);
            }
            E.time_relaxation_damage = dmax(o.time_relaxation_damage * o.deltaT_relaxation_damage / deltaT, ddt);
        } else {
            E.time_relaxation_damage = 1e36;
        }
    }

    // ---- diagnostics (FE.cpp:5903-5984) ----
    A.Qa[i] = Fi.Qia * old_conc + Fy.Qia * old_conc_young + Qow * old_ow_fraction;
    A.Qo[i] = Qio_mean + Qow_mean;
    A.Qnosun[i] = Qio_mean + Qnosun_ow;
    A.Qassim[i] = Qassm;
    A.delS[i] = delsss * phys::rhow * mld * days_in_sec / o.dtime_step;
    A.fwflux_ice[i] = -1. / P.u_ddt * ((1. - 1e-3 * si_eff) * phys::rhoi * del_vi + phys::rhos * del_vs_mlt);
    A.fwflux[i] = A.fwflux_ice[i] - emp;
    A.brine[i] = -1e-3 * si_eff * phys::rhoi * del_vi / P.u_ddt;
    A.evap[i] = evap * (1. - old_conc - old_conc_young);
    A.rain[i] = rain;
    A.vice_melt[i] = del_vi * days_in_sec / P.u_ddt;
    A.del_vi_young[i] = del_vi_young * days_in_sec / P.u_ddt;
    A.del_hi[i] = del_hi * days_in_sec / P.u_ddt;
    A.del_hi_young[i] = del_hi_young * days_in_sec / P.u_ddt;
    A.newice[i] = newice_stored * days_in_sec / P.u_ddt;
    A.mlt_top[i] = mlt_vi_top * days_in_sec / P.u_ddt;
    A.mlt_bot[i] = mlt_vi_bot * days_in_sec / P.u_ddt;
    A.snow2ice[i] = snow2ice * days_in_sec / P.u_ddt;


    // ---- age / multi-year-ice tracers (FE.cpp:5986-6131) ----
    E.fyi_fraction = A.fyi_fraction[i]; E.age_det = A.age_det[i]; E.age = A.age[i]; E.freeze_onset = A.freeze_onset[i];
    E.conc_myi = A.conc_myi[i]; E.thick_myi = A.thick_myi[i];
    double del_vi_rplnt_myi = 0., del_ci_rplnt_myi = 0., del_vi_mlt_myi = 0., del_ci_mlt_myi = 0.;
    if (E.conc < phys::cmin || E.thick < E.conc * phys::hmin) {
        E.fyi_fraction = 0.;
        E.age_det = 0.;
        E.age = 0.;
        E.thick_myi = 0.;
        E.conc_myi = 0.;
        E.freeze_days = 0.;
        E.freeze_onset = 1.;
    } else {
        if (P.is_0915 && P.midnight) {
            E.fyi_fraction = 0.;
        } else {
            double conc_fyi = E.fyi_fraction + del_c;
            E.fyi_fraction = dmax(0., dmin(1., conc_fyi));
        }
        double w_age = old_conc <= 0 ? 0. : dmin(old_conc / E.conc, 1.);
        E.age_det = w_age * (E.age_det + dt) + dmax((1 - w_age) * dt, 0.);
        w_age = old_vol <= 0 ? 0. : dmin(old_vol / E.thick, 1.);
        E.age = w_age * (E.age + dt) + dmax((1 - w_age) * dt, 0.);

        bool reset_myi = false;
        if (o.reset_by_date) {
            if (P.is_reset_date && P.midnight) reset_myi = true;
        } else {
            if (E.freeze_days >= o.freeze_days_threshold) {
                if (E.freeze_onset <= 0.5) {
                    reset_myi = true;
                    E.freeze_onset = 1.;
                }
            }
        }
        if (P.is_0801 && P.midnight) {
            E.freeze_onset = 0.;
            double ctot = E.conc;
            if (young) ctot += E.conc_young;
            if (ctot == 0.) E.freeze_onset = 1.;
            double conc_summer = E.conc;
            double thick_summer = E.thick;
            if (young && use_young_ice_in_myi_reset) {
                conc_summer += E.conc_young;
                thick_summer += E.h_young;
            }
            E.conc_summer = dmax(0., dmin(1., conc_summer));
            E.thick_summer = dmax(0., thick_summer);
        }
        E.freeze_onset = round(E.freeze_onset);

        double old_conc_myi = E.conc_myi;
        double old_thick_myi = E.thick_myi;
        double c_myi_max = E.conc;
        double v_myi_max = E.thick;
        if (young && use_young_ice_in_myi_reset) {
            c_myi_max += E.conc_young;
            v_myi_max += E.h_young;
        }
        if (reset_myi) {
            if (!o.reset_by_date) {
                double c_myi_reset = dmax(E.conc_summer, E.conc_myi);
                double v_myi_reset = dmax(E.thick_summer, E.thick_myi);
                E.conc_myi = dmin(c_myi_max, c_myi_reset);
                E.thick_myi = dmin(v_myi_max, v_myi_reset);
            } else {
                E.conc_myi = c_myi_max;
                E.thick_myi = v_myi_max;
            }
            E.conc_myi = dmax(0., dmin(1., E.conc_myi));
            E.thick_myi = dmax(0., E.thick_myi);
            del_ci_rplnt_myi = E.conc_myi - old_conc_myi;
            del_vi_rplnt_myi = E.thick_myi - old_thick_myi;
        } else {
            if ((E.thick < old_vol) && (old_conc > 0) && (old_vol > 0)) {
                if (o.equal_melting) {
                    double const del_c_ratio = dmin(E.conc / old_conc, 1.);
                    double const del_v_ratio = dmin(E.thick / old_vol, 1.);
                    del_ci_mlt_myi = dmin(0., E.conc_myi * (del_c_ratio - 1.));
                    del_vi_mlt_myi = dmin(0., E.thick_myi * (del_v_ratio - 1.));
                }
                E.conc_myi = dmax(0., dmin(c_myi_max, E.conc_myi + del_ci_mlt_myi));
                E.thick_myi = dmax(0., dmin(v_myi_max, E.thick_myi + del_vi_mlt_myi));
                del_ci_mlt_myi = E.conc_myi - old_conc_myi;
                del_vi_mlt_myi = E.thick_myi - old_thick_myi;
            }
        }
    }
    A.del_ci_mlt_myi[i] = del_ci_mlt_myi * days_in_sec / P.u_ddt;
    A.del_vi_mlt_myi[i] = del_vi_mlt_myi * days_in_sec / P.u_ddt;
    A.del_ci_rplnt_myi[i] = del_ci_rplnt_myi * days_in_sec / P.u_ddt;
    A.del_vi_rplnt_myi[i] = del_vi_rplnt_myi * days_in_sec / P.u_ddt;
}

// thermo() for element i: the fields the options make it read are loaded once, the ones it can change are stored once.
// Which forcing variable feeds a quantity follows ExternalData::isInitialized() exactly as in the reference
// (specificHumidity 4980-4986, incomingLongwave 6378, the snowfall cascade 5333-5340, M_mld 5346).
NSX_HD void thermo_element(Params const& P, Arrays const& A, int i)
{
    NsxThermoParams const& o = P.o;
    bool const young = o.ice_cat_young != 0, winton = o.thermo_type == 1, ponds = o.use_meltponds != 0;
    Elem E;
    // forcing
    E.tair = A.tair[i]; E.mslp = A.mslp[i]; E.Qsw_in = A.Qsw_in[i]; E.precip = A.precip[i];
    E.sphuma = o.have_sphuma ? A.sphuma[i] : 0.;
    E.mixrat = (!o.have_sphuma && o.have_mixrat) ? A.mixrat[i] : 0.;
    E.dair = (!o.have_sphuma && !o.have_mixrat) ? A.dair[i] : 0.;
    E.Qlw_in = o.have_Qlw_in ? A.Qlw_in[i] : 0.;
    E.tcc = o.have_Qlw_in ? 0. : A.tcc[i];
    E.snowfr = o.have_snowfr ? A.snowfr[i] : 0.;
    E.snowfall = (!o.have_snowfr && o.have_snowfall) ? A.snowfall[i] : 0.;
    E.mld = o.have_mld ? A.mld[i] : 0.;
    E.ocean_temp = o.ocean_constant ? 0. : A.ocean_temp[i];
    E.ocean_salt = o.ocean_constant ? 0. : A.ocean_salt[i];
    E.conc_upd = o.use_assim_flux ? A.conc_upd[i] : 0.;     // only read under use_assim_flux (FE.cpp:5428)
    // ice state
    E.conc = A.conc[i]; E.thick = A.thick[i]; E.snow_thick = A.snow_thick[i]; E.drag_ui = A.drag_ui[i];
    E.time_relaxation_damage = 0.;                          // written, never read
    E.conc_young = young ? A.conc_young[i] : 0.; E.h_young = young ? A.h_young[i] : 0.; E.hs_young = young ? A.hs_young[i] : 0.;
    E.drag_ui_young = young ? A.drag_ui_young[i] : 0.;
    // slab ocean, ice temperatures, tracers
    E.sst = A.sst[i]; E.sss = A.sss[i]; E.tice0 = A.tice0[i];
    E.tice1 = winton ? A.tice1[i] : 0.; E.tice2 = winton ? A.tice2[i] : 0.;
    E.tsurf_young = young ? A.tsurf_young[i] : 0.; E.drag_ti_young = young ? A.drag_ti_young[i] : 0.;
    E.drag_ti = A.drag_ti[i];
    E.conc_summer = E.thick_summer = 0.;                    // the tracer state is loaded inside thermo_core() where it is first used
    E.pond_fraction = A.pond_fraction[i];
    E.pond_volume = ponds ? A.pond_volume[i] : 0.;
    E.lid_volume = (ponds || E.pond_fraction > 0.) ? A.lid_volume[i] : 0.;       // FE.cpp:6327 reads it only over a pond

    thermo_core(P, A, i, E);

    A.conc[i] = E.conc; A.thick[i] = E.thick; A.snow_thick[i] = E.snow_thick; A.ridge_ratio[i] = E.ridge_ratio;
    A.conc_myi[i] = E.conc_myi; A.thick_myi[i] = E.thick_myi; A.drag_ui[i] = E.drag_ui;
    if (o.temp_dep_healing) A.time_relaxation_damage[i] = E.time_relaxation_damage;
    if (young) {
        A.conc_young[i] = E.conc_young; A.h_young[i] = E.h_young; A.hs_young[i] = E.hs_young; A.drag_ui_young[i] = E.drag_ui_young;
        A.tsurf_young[i] = E.tsurf_young; A.drag_ti_young[i] = E.drag_ti_young;
    }
    A.sst[i] = E.sst; A.sss[i] = E.sss; A.tice0[i] = E.tice0;
    if (winton) { A.tice1[i] = E.tice1; A.tice2[i] = E.tice2; }
    A.del_vi_tend[i] = E.del_vi_tend; A.freeze_days[i] = E.freeze_days; A.freeze_onset[i] = E.freeze_onset;
    A.fyi_fraction[i] = E.fyi_fraction; A.age_det[i] = E.age_det; A.age[i] = E.age; A.drag_ti[i] = E.drag_ti;
    // M_conc_summer / M_thick_summer change on the last step of a day and at the 1 August midnight only (FE.cpp:5680-5697, 6059-6077)
    if (P.step_in_day == P.num_steps_in_day || (P.is_0801 && P.midnight)) { A.conc_summer[i] = E.conc_summer; A.thick_summer[i] = E.thick_summer; }
    if (ponds) { A.pond_volume[i] = E.pond_volume; A.lid_volume[i] = E.lid_volume; A.pond_fraction[i] = E.pond_fraction; }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: options -> Params (the scalars thermo() and IABulkFluxes derive before their loops)
// ---------------------------------------------------------------------------------------------------------------------
// nextsim time (decimal days since 1900-01-01 00:00) -> month, day of datenumToString(t, "%m%d") (core/include/date.hpp:87-120):
// the date is 1900-01-01 + static_cast<long>(t) days, the time of day is rounded to milliseconds, and a time of day that rounds
// to 24:00:00.000 carries into the next date (boost::posix_time::ptime(date, time_duration)).
inline void month_day(double datenum, int& month, int& day)
{
    long days = static_cast<long>(datenum);
    double const frac = datenum - std::floor(datenum);
    if (static_cast<long>(std::floor(frac * 24.0 * 60.0 * 60.0 * 1000.0 + 0.5)) >= 86400000L) ++days;
    long const z = days + 693901L;                          // days since 0000-03-01 (1970-01-01 is day 719468, 1900-01-01 is 25567 earlier)
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
    auto ud = [](double c) { return UDiv{c, 1. / c}; };
    P.u_ddt = ud(P.ddt); P.u_rhos = ud(phys::rhos); P.u_rhow = ud(phys::rhow); P.u_rhoi = ud(phys::rhoi);
    P.u_qi = ud(P.qi); P.u_qs = ud(P.qs); P.u_ki = ud(phys::ki); P.u_h_young_min = ud(o.h_young_min);
    P.u_2ddt = ud(2 * P.ddt); P.u_Crho = ud(phys::C * phys::rhoi);
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
