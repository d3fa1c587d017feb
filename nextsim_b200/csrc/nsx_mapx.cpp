// nsx_mapx.cpp -- node latitudes for the path, host side.
//
// FiniteElement::explicitSolve() calls M_mesh.lat() (FE.cpp:10351) for the Coriolis term and the sign of the turning
// angle; GmshMesh::lat() (core/src/gmshmesh.cpp:1800-1824) is init_mapx(mesh.mppfile) + inverse_mapx() per node, with
// the in-tree mapx library (contrib/mapx).  The meshes of the reference use the two "Polar Stereographic [Ellipsoid]"
// files mesh/NpsNextsim.mpp and mesh/NpsASR.mpp in mapx's positional (pre-keyword) format; this file implements exactly
// that subset: the positional .mpp reader (mapx.c:753-880), the rotation / origin handling of reinit_mapx and
// inverse_mapx (mapx.c:971-1063, 1129-1146) and the inverse polar stereographic maps (polar_stereographic.c).
// tests/test_ref_mapx_cpu.py pins it against the reference's own mapx built unmodified into oracle/_ref.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nsx.h"

namespace {

thread_local std::string g_mapx_err;
const double MAPX_PI = 3.141592653589793;          // mapx.h:47
const double MAPX_RE_KM = 6371.228;                // mapx.h:55
const double MAPX_ECC = 0.082271673;               // mapx.h:58

struct Projection {
    bool ellipsoid = false;
    double lat0 = 0, lon0 = 0, lat1 = 999, rotation = 0, scale = 1, center_lat = 0, center_lon = 0;
    double equatorial_radius = MAPX_RE_KM, eccentricity = 0;
    // derived
    double Rg = 0, e2 = 0, e4 = 0, e6 = 0, e8 = 0, sin_phi1 = 0, m1 = 0, t1 = 0;
    double T00 = 1, T01 = 0, T10 = 0, T11 = 1, u0 = 0, v0 = 0;
};

inline double rad(double t) { return t * MAPX_PI / 180; }
inline double deg(double t) { return t * 180 / MAPX_PI; }

std::string squeeze_upper(std::string const& s)
{
    std::string o;
    for (char c : s) if (c != ' ' && c != '\t' && c != '_' && c != '-' && c != '\r') o += (char)std::toupper((unsigned char)c);
    return o;
}

// positional .mpp: name / lat0 lon0 [lat1 [lon1]] / rotation / scale / center lat lon / 5 display lines /
// [equatorial radius] / [eccentricity]; text after the numbers is a comment
Projection read_mpp(const char* path)
{
    std::ifstream in(path);
    if (!in) throw std::invalid_argument(std::string("mapx: cannot open ") + path);
    std::vector<std::string> lines;
    for (std::string l; std::getline(in, l);) {
        if (l.find_first_not_of(" \t\r") == std::string::npos && lines.size() >= 10) continue;   // trailing blank lines
        lines.push_back(l);
    }
    if (lines.size() < 10) throw std::runtime_error("mapx: map projection parameters file is too short");
    Projection P;
    std::string const name = squeeze_upper(lines[0]);
    if (name == "POLARSTEREOGRAPHICELLIPSOID") P.ellipsoid = true;
    else if (name != "POLARSTEREOGRAPHIC")
        throw std::runtime_error("mapx: projection '" + lines[0] + "' is not supported by this library (polar stereographic only)");
    auto nums = [&](size_t i, double* v, int n) { std::istringstream s(lines[i]); int k = 0; while (k < n && (s >> v[k])) ++k; return k; };
    double v[4];
    int k = nums(1, v, 4);
    P.lat0 = k >= 1 ? v[0] : 0.0; P.lon0 = k >= 2 ? v[1] : 0.0; P.lat1 = k >= 3 ? v[2] : 999;
    P.rotation = nums(2, v, 1) >= 1 ? v[0] : 0.0;
    P.scale = nums(3, v, 1) >= 1 ? v[0] : 1.0;
    k = nums(4, v, 2);
    P.center_lat = k >= 1 ? v[0] : 0.0; P.center_lon = k >= 2 ? v[1] : 0.0;
    P.equatorial_radius = MAPX_RE_KM;
    P.eccentricity = P.ellipsoid ? MAPX_ECC : 0.0;
    if (lines.size() > 10 && nums(10, v, 1) >= 1) P.equatorial_radius = v[0];
    if (lines.size() > 11 && nums(11, v, 1) >= 1) P.eccentricity = v[0];
    if (P.lat0 != 90.0 && P.lat0 != -90.0) throw std::runtime_error("mapx: only polar aspects allowed");
    return P;
}

void forward(Projection const& P, double lat, double lon, double& x, double& y)
{
    bool const north = (P.lat0 == 90.0);
    if (P.ellipsoid) {
        double const phi = north ? rad(lat) : rad(-lat);
        double const lam = north ? rad(lon - P.lon0) : rad(-lon + P.lon0);
        double const sp = std::sin(phi);
        double const t = std::sqrt((1.0 - sp) / (1.0 + sp) *
                                   std::pow((1.0 + P.eccentricity * sp) / (1.0 - P.eccentricity * sp), P.eccentricity));
        double rho;
        if (P.lat1 != 90.0 && P.lat1 != -90.0) rho = P.Rg * P.m1 * t / P.t1;
        else rho = 2 * P.Rg * P.scale * t / std::sqrt(std::pow(1 + P.eccentricity, 1 + P.eccentricity) *
                                                      std::pow(1 - P.eccentricity, 1 - P.eccentricity));
        x = rho * std::sin(lam); y = -rho * std::cos(lam);
        if (!north) { x = -x; y = -y; }
    } else {
        double const phi = rad(lat), lam = rad(lon - P.lon0);
        if (north) {
            double const rho = P.Rg * std::cos(phi) * (1 + P.sin_phi1) / (1 + std::sin(phi));
            x = rho * std::sin(lam); y = -rho * std::cos(lam);
        } else {
            double const rho = P.Rg * std::cos(phi) * (1 - P.sin_phi1) / (1 - std::sin(phi));
            x = rho * std::sin(lam); y = rho * std::cos(lam);
        }
    }
}

void init(Projection& P)
{
    P.e2 = P.eccentricity * P.eccentricity; P.e4 = P.e2 * P.e2; P.e6 = P.e4 * P.e2; P.e8 = P.e4 * P.e4;
    P.Rg = P.equatorial_radius / P.scale;
    if (P.lat1 == 999) P.lat1 = P.lat0;
    double const s = (P.lat0 == 90.0) ? 1.0 : -1.0;
    if (P.ellipsoid) {
        double const cos_phi1 = std::cos(rad(s * P.lat1));
        P.sin_phi1 = std::sin(rad(s * P.lat1));
        P.m1 = cos_phi1 / std::sqrt(1 - (P.e2 * P.sin_phi1 * P.sin_phi1));
        double const num = 1 - P.eccentricity * P.sin_phi1, den = 1 + P.eccentricity * P.sin_phi1;
        P.t1 = std::tan(MAPX_PI / 4 - rad(s * P.lat1) / 2) / std::pow(num / den, P.eccentricity / 2);
    } else {
        P.sin_phi1 = std::sin(rad(P.lat1));
    }
    double const theta = rad(P.rotation);
    P.T00 = std::cos(theta); P.T01 = std::sin(theta); P.T10 = -std::sin(theta); P.T11 = std::cos(theta);
    double x0, y0;
    forward(P, P.center_lat, P.center_lon, x0, y0);          // the map origin is the projected centre
    P.u0 = P.T00 * x0 + P.T01 * y0;
    P.v0 = P.T10 * x0 + P.T11 * y0;
}

void inverse(Projection const& P, double u, double v, double& lat, double& lon)
{
    u += P.u0; v += P.v0;
    double const x = P.T00 * u - P.T01 * v, y = -P.T10 * u + P.T11 * v;
    double const rho = std::sqrt(x * x + y * y);
    bool const north = (P.lat0 == 90.0);
    double phi, lam;
    if (P.ellipsoid) {
        double t;
        if (P.lat1 == 90.0 || P.lat1 == -90.0)
            t = rho * std::sqrt(std::pow(1 + P.eccentricity, 1 + P.eccentricity) * std::pow(1 - P.eccentricity, 1 - P.eccentricity)) /
                (2 * P.Rg * P.scale);
        else
            t = (rho * P.t1) / (P.Rg * P.m1);
        double const chi = MAPX_PI / 2.0 - 2.0 * std::atan(t);
        double const s2 = std::sin(2.0 * chi), s4 = std::sin(4.0 * chi), s6 = std::sin(6.0 * chi);
        phi = chi + (s2 * P.e2 / 2.0) + (s2 * 5.0 * P.e4 / 24.0) + (s2 * P.e6 / 12.0) + (s2 * 13.0 * P.e8 / 360.0)
            + (s4 * 7.0 * P.e4 / 48.0) + (s4 * 29.0 * P.e6 / 240.0) + (s4 * 811.0 * P.e8 / 11520.0)
            + (s6 * 7.0 * P.e6 / 120.0) + (s6 * 81.0 * P.e8 / 1120.0) + (std::sin(8.0 * chi) * 4279.0 * P.e8 / 161280.0);
        if (north) { lat = deg(phi); lam = std::atan2(x, -y); lon = deg(lam) + P.lon0; }
        else { lat = -deg(phi); lam = std::atan2(-x, y); lon = -deg(lam) + P.lon0; }
    } else {
        if (north) { double const c = 2 * std::atan2(rho, P.Rg * (1 + P.sin_phi1)); phi = std::asin(std::cos(c)); lam = std::atan2(x, -y); }
        else { double const c = 2 * std::atan2(rho, P.Rg * (1 - P.sin_phi1)); phi = std::asin(-std::cos(c)); lam = std::atan2(x, y); }
        lat = deg(phi); lon = deg(lam) + P.lon0;
    }
    while (lon < -180) lon += 360;
    while (lon > 180) lon -= 360;
}

} // namespace

extern "C" const char* nsx_mapx_last_error(void) { return g_mapx_err.c_str(); }

// GmshMesh::lat() / lon(): inverse map of n points given in map units (the mesh coordinates)
extern "C" int nsx_mapx_latlon(const char* mppfile, int n, const double* x, const double* y, double* lat, double* lon)
{
    try {
        if (!mppfile || !x || !y || (!lat && !lon) || n < 0) throw std::invalid_argument("nsx_mapx_latlon: bad argument");
        Projection P = read_mpp(mppfile);
        init(P);
        for (int i = 0; i < n; ++i) {
            double la, lo;
            inverse(P, x[i], y[i], la, lo);
            if (lat) lat[i] = la;
            if (lon) lon[i] = lo;
        }
    } catch (std::exception const& e) { g_mapx_err = e.what(); return 2; }
    return 0;
}
