// thermo_oracle.cpp -- CPU side of the thermo() parity check (TEST INFRASTRUCTURE ONLY, oracle/).
//
// thermo() is one element-wise function, so the restatement of FE.cpp:4966-6962 exists ONCE, as the host+device function
// nsx::thermo::thermo_element() of nextsim_b200/csrc/nsx_thermo.cuh; this file compiles that function for the host (g++,
// -ffp-contract=off like the rest of the oracle) and loops it over the elements.  It is the checker of the device build
// of the same text: the two differ by the device libm and by the marked reciprocal / integer-power shortcuts of the device
// build (<= 1-2 ulp each, header of nsx_thermo.cuh).  What pins the TEXT to the reference is not this file but
// oracle/ref_fe -- the reference's own thermo(), OWBulkFluxes(), IABulkFluxes(), thermoWinton(), thermoIce0(), ... bodies cut
// from /root/reference at build time -- against which tests/test_thermo_cpu.py holds this build BIT FOR BIT over every
// option branch, and from which tests/golden/thermo/*.npz were generated.  PARITY PINNED.
#include <cstring>
#include <string>
#include <vector>

#include "../nextsim_b200/csrc/nsx_thermo.cuh"

namespace {
std::string g_names;
std::string g_err;
}

extern "C" {

// space-separated reference member names orc_thermo() accepts, in a fixed order
const char* orc_thermo_field_names()
{
    if (g_names.empty()) {
#define X(n) g_names += "M_" #n " ";
        NSX_THERMO_FORCING(X)
        NSX_THERMO_ICE(X)
        NSX_THERMO_STATE(X)
#undef X
        g_names += "D_pond_fraction ";
#define X(n) g_names += "D_" #n " ";
        NSX_THERMO_DIAG(X)
#undef X
        g_names.pop_back();
    }
    return g_names.c_str();
}
const char* orc_thermo_last_error() { return g_err.c_str(); }

void orc_thermo_params_defaults(NsxThermoParams* p) { nsx::thermo::params_defaults(*p); }
int orc_thermo_params_size() { return (int)sizeof(NsxThermoParams); }
// month * 100 + day of datenumToString(t, "%m%d") as the product decodes it (nsx::thermo::month_day)
int orc_thermo_month_day(double datenum) { int m, d; nsx::thermo::month_day(datenum, m, d); return 100 * m + d; }

// tri0: [3*ne] 0-based node ids, element-major; wind / VT / ocean: [2*nn]; names[k] -> ptrs[k] ([ne] doubles, updated in place)
int orc_thermo(const NsxThermoParams* o, int dt, double current_time, int ne, int nn, const int* tri0, const double* wind,
               const double* VT, const double* ocean, int nfields, const char** names, double** ptrs)
{
    using namespace nsx::thermo;
    if (const char* e = validate(*o, dt)) { g_err = e; return 2; }
    Arrays A;
    std::memset(&A, 0, sizeof A);
    A.ne = ne; A.nn = nn;
    std::vector<int> e0(ne), e1(ne), e2(ne);
    for (int i = 0; i < ne; ++i) { e0[i] = tri0[3 * i]; e1[i] = tri0[3 * i + 1]; e2[i] = tri0[3 * i + 2]; }
    A.en0 = e0.data(); A.en1 = e1.data(); A.en2 = e2.data();
    A.wind = wind; A.VT = VT; A.ocean = ocean;
    for (int k = 0; k < nfields; ++k) {
        double** s = field_slot(A, names[k]);
        if (!s) { g_err = std::string("unknown field ") + names[k]; return 2; }
        *s = ptrs[k];
    }
    // every field must be present (the caller passes zeros for what the configuration does not use)
    std::string all = orc_thermo_field_names();
    size_t pos = 0;
    while (pos < all.size()) {
        size_t e = all.find(' ', pos);
        if (e == std::string::npos) e = all.size();
        std::string n = all.substr(pos, e - pos);
        if (!*field_slot(A, n.c_str())) { g_err = "missing field " + n; return 2; }
        pos = e + 1;
    }
    Params P = make_params(*o, dt, current_time);
    for (int i = 0; i < ne; ++i) thermo_element(P, A, i);
    return 0;
}

}  // extern "C"
