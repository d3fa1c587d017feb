// nsx_cfg.cpp -- host-side option handling for the hot path: defaults of model/options.cpp and a small
// reader for nextsim.cfg INI files (the reference uses boost::program_options, environment.cpp:43-72,
// which is not available here).  Only the keys the path consumes are interpreted;
// the other keys of the sections we look at are accepted and ignored, unknown [dynamics] keys are an
// error like in the reference (parse_config_file(..., allow_unregistered=false), environment.cpp:68).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <string>

#include "../../include/nsx.h"

static std::string g_cfg_err;
extern "C" const char* nsx_cfg_last_error() { return g_cfg_err.c_str(); }

extern "C" void nsx_params_defaults(NsxDynParams* p)
{
    std::memset(p, 0, sizeof(*p));
    p->dynamics_type = NSX_DYN_BBM;             // options.cpp:111
    p->basal_stress_type = NSX_BASAL_LEMIEUX;   // options.cpp:109
    p->newice_type = 4;                         // options.cpp:397
    p->ice_cat_type = NSX_ICECAT_YOUNG_ICE;     // FE.cpp:1212-1215
    p->substeps = 120;                          // options.cpp:363
    p->equal_ridging = 0;                       // options.cpp:547
    p->use_young_ice_in_myi_reset = 1;          // options.cpp:545
    p->use_coriolis = 1;                        // options.cpp:346
    p->dtime_step = 200.;                       // options.cpp:43
    p->ocean_turning_angle_rad = (3.14159265358979323846 / 180.) * 25.;   // options.cpp:347, FE.cpp:1172
    p->min_h = 0.05; p->min_c = 0.01;           // options.cpp:325-326
    p->young = 5.9605e+08;                      // options.cpp:313
    p->nu0 = 1. / 3.;                           // options.cpp:318
    p->tan_phi = 0.7;                           // options.cpp:319
    p->compr_strength = 1e10;                   // options.cpp:320 (unscaled; host multiplies by scale_coef)
    p->compaction_param = -20.;                 // options.cpp:321
    p->undamaged_time_relaxation_sigma = 1e7;   // options.cpp:331
    p->exponent_relaxation_sigma = 5.;          // options.cpp:333
    p->compression_factor = 10e3;               // options.cpp:359
    p->exponent_compression_factor = 1.5;       // options.cpp:358
    p->quad_drag_coef_water = 0.0055;           // options.cpp:342
    p->evp_e = 2.; p->evp_Pstar = 27.5e3; p->evp_C = 20.; p->evp_dmin = 1e-9;   // options.cpp:365-372
    p->mevp_alpha = 500.; p->mevp_beta = 500.;  // options.cpp:375-376
    p->basal_k1 = 10.; p->basal_k2 = 15.; p->basal_Cb = 20.; p->basal_u0 = 5e-5;   // options.cpp:350-353
    p->C_lab = 2.0e6;                           // options.cpp:317
    p->alea_factor = 0.;                        // options.cpp:311
    p->time_relaxation_damage_days = 25.;       // options.cpp:329
}

static std::string trim(std::string s)
{
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}

static bool to_bool(std::string v, bool& out)
{
    std::transform(v.begin(), v.end(), v.begin(), ::tolower);
    if (v == "true" || v == "1" || v == "yes" || v == "on") { out = true; return true; }
    if (v == "false" || v == "0" || v == "no" || v == "off") { out = false; return true; }
    return false;
}

extern "C" int nsx_params_from_cfg(const char* path, NsxDynParams* p)
{
    nsx_params_defaults(p);
    std::ifstream in(path);
    if (!in) { g_cfg_err = std::string("cannot open ") + path; return 1; }
    // [dynamics] keys the reference declares but the hot path does not read (options.cpp:309-376)
    static const std::set<std::string> ignored_dynamics = {
        "use_temperature_dependent_healing", "deltaT_relaxation_damage", "ERA5_quad_drag_coef_air",
        "ECMWF_quad_drag_coef_air", "ASR_quad_drag_coef_air", "CFSR_quad_drag_coef_air", "lin_drag_coef_air",
        "lin_drag_coef_water", "Lemieux_basal_u_crit"};
    std::string line, section;
    bool use_coriolis = true;
    double turning_deg = 25.;
    int lineno = 0;
    while (std::getline(in, line)) {
        ++lineno;
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        line = trim(line);
        if (line.empty()) continue;
        if (line.front() == '[') {
            size_t close = line.find(']');
            if (close == std::string::npos) { g_cfg_err = "line " + std::to_string(lineno) + ": bad section"; return 2; }
            section = trim(line.substr(1, close - 1));
            continue;
        }
        size_t eq = line.find('=');
        if (eq == std::string::npos) { g_cfg_err = "line " + std::to_string(lineno) + ": expected key=value"; return 2; }
        std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        auto num = [&](double& dst) -> bool {
            char* end = nullptr;
            double v = std::strtod(val.c_str(), &end);
            if (end == val.c_str() || *end != 0) { g_cfg_err = section + "." + key + ": not a number: " + val; return false; }
            dst = v;
            return true;
        };
        // po::value<int> (options.cpp:43, 363, 397) rejects anything that is not an integer literal
        auto integer = [&](int& dst) -> bool {
            char* end = nullptr;
            long v = std::strtol(val.c_str(), &end, 10);
            if (end == val.c_str() || *end != 0) { g_cfg_err = section + "." + key + ": not an integer: " + val; return false; }
            dst = (int)v;
            return true;
        };
        auto boolean = [&](int& dst) -> bool {
            bool b;
            if (!to_bool(val, b)) { g_cfg_err = section + "." + key + ": not a bool: " + val; return false; }
            dst = b;
            return true;
        };
        bool ok = true;
        if (section == "dynamics") {
            if (key == "substeps") ok = integer(p->substeps);
            else if (key == "young") ok = num(p->young);
            else if (key == "nu0") ok = num(p->nu0);
            else if (key == "C_lab") ok = num(p->C_lab);
            else if (key == "alea_factor") ok = num(p->alea_factor);
            else if (key == "tan_phi") ok = num(p->tan_phi);
            else if (key == "compr_strength") ok = num(p->compr_strength);
            else if (key == "compaction_param") ok = num(p->compaction_param);
            else if (key == "min_h") ok = num(p->min_h);
            else if (key == "min_c") ok = num(p->min_c);
            else if (key == "time_relaxation_damage") ok = num(p->time_relaxation_damage_days);
            else if (key == "undamaged_time_relaxation_sigma") ok = num(p->undamaged_time_relaxation_sigma);
            else if (key == "exponent_relaxation_sigma") ok = num(p->exponent_relaxation_sigma);
            else if (key == "compression_factor") ok = num(p->compression_factor);
            else if (key == "exponent_compression_factor") ok = num(p->exponent_compression_factor);
            else if (key == "quad_drag_coef_water") ok = num(p->quad_drag_coef_water);
            else if (key == "use_coriolis") { int b; ok = boolean(b); use_coriolis = b; }
            else if (key == "oceanic_turning_angle") ok = num(turning_deg);
            else if (key == "Lemieux_basal_k1") ok = num(p->basal_k1);
            else if (key == "Lemieux_basal_k2") ok = num(p->basal_k2);
            else if (key == "Lemieux_basal_Cb") ok = num(p->basal_Cb);
            else if (key == "Lemieux_basal_u_0") ok = num(p->basal_u0);
            else if (key == "evp.e") ok = num(p->evp_e);
            else if (key == "evp.Pstar") ok = num(p->evp_Pstar);
            else if (key == "evp.C") ok = num(p->evp_C);
            else if (key == "evp.dmin") ok = num(p->evp_dmin);
            else if (key == "mevp.alpha") ok = num(p->mevp_alpha);
            else if (key == "mevp.beta") ok = num(p->mevp_beta);
            else if (!ignored_dynamics.count(key)) { g_cfg_err = "unrecognised option 'dynamics." + key + "'"; return 3; }
        } else if (section == "setup") {
            if (key == "dynamics-type") {      // options.cpp:111, FE.cpp str2dynamics map
                if (val == "bbm") p->dynamics_type = NSX_DYN_BBM;
                else if (val == "evp") p->dynamics_type = NSX_DYN_EVP;
                else if (val == "mevp") p->dynamics_type = NSX_DYN_MEVP;
                else { g_cfg_err = "setup.dynamics-type=" + val + " is outside the accelerated path (bbm|evp|mevp)"; return 4; }
            } else if (key == "basal_stress-type") {
                if (val == "none") p->basal_stress_type = NSX_BASAL_NONE;
                else if (val == "lemieux") p->basal_stress_type = NSX_BASAL_LEMIEUX;
                else { g_cfg_err = "invalid option for setup.basal_stress-type: " + val; return 4; }
            }
        } else if (section == "simul") {
            if (key == "timestep") { int ts = 0; ok = integer(ts); if (ok) p->dtime_step = ts; }     // po::value<int>, options.cpp:43
        } else if (section == "thermo") {
            if (key == "newice_type") ok = integer(p->newice_type);
        } else if (section == "age") {
            if (key == "equal_ridging") ok = boolean(p->equal_ridging);
            else if (key == "include_young_ice") ok = boolean(p->use_young_ice_in_myi_reset);
        }
        if (!ok) return 2;
    }
    p->use_coriolis = use_coriolis;
    p->ocean_turning_angle_rad = use_coriolis ? (3.14159265358979323846 / 180.) * turning_deg : 0.;   // FE.cpp:1167-1172
    p->ice_cat_type = (p->newice_type == 4) ? NSX_ICECAT_YOUNG_ICE : NSX_ICECAT_CLASSIC;               // FE.cpp:1212-1215
    return 0;
}
