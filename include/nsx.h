/* nsx.h -- C ABI of the B200-native neXtSIM explicit momentum / rheology solver (libnsx.so).
 *
 * This is the drop-in boundary for ONE hot path of nansencenter/nextsim:
 *   FiniteElement::explicitSolve()      model/finiteelement.cpp:10182-10643
 *   FiniteElement::updateSigmaDamage()  model/finiteelement.cpp:4137-4260   (BBM)
 *   FiniteElement::updateSigmaEVP/MEVP  model/finiteelement.cpp:10649-10726
 *   FiniteElement::updateGhosts()       model/finiteelement.cpp:13963-13996
 *   FiniteElement::update()             model/finiteelement.cpp:3919-4132
 * The reference has no FFI layer: the path is a set of member functions of FiniteElement working on
 * its std::vector<double> members (declared model/finiteelement.hpp:166-169, 298-303, 521-522).
 * Each entry point below names the member function or member set it replaces.  Host code keeps
 * ownership of every array; the handle owns device memory only.  Plain pointers and sizes, no C++
 * or torch types.  All functions return 0 on success, non-zero on error (see nsx_last_error);
 * the host shim turns a non-zero status into the std::runtime_error the reference would throw.
 *
 * Layout conventions (identical to the reference):
 *   - element arrays: [num_elements], owned elements first, ghost elements after
 *   - nodal scalars:  [num_nodes], owned nodes first (local_ndof), ghost nodes after
 *   - nodal vectors:  [2*num_nodes] split storage  u[0..N) | v[0..N)   (FE.cpp:10331-10332)
 *   - indices[]: 1-based local node ids, 3 per element (GmshMesh::indexTr(), gmshmesh.cpp:1544)
 */
#ifndef NSX_H
#define NSX_H

#ifdef __cplusplus
extern "C" {
#endif

#define NSX_VERSION 1

typedef struct nsx_solver* nsx_handle;

/* setup::DynamicsType (model/enums.hpp:142-149) */
enum { NSX_DYN_BBM = 0, NSX_DYN_EVP = 3, NSX_DYN_MEVP = 4 };
/* setup::BasalStressType (enums.hpp:86-90), setup::IceCategoryType (enums.hpp:92-97) */
enum { NSX_BASAL_NONE = 0, NSX_BASAL_LEMIEUX = 1 };
enum { NSX_ICECAT_CLASSIC = 0, NSX_ICECAT_YOUNG_ICE = 1 };

/* Options read by the path; names and defaults follow model/options.cpp:43,109,111,309-376,397,545-547.
 * Filled by nsx_params_defaults()/nsx_params_from_cfg() or by the host from its own vm[...] map. */
typedef struct NsxDynParams {
    int dynamics_type;              /* setup.dynamics-type           (bbm)      */
    int basal_stress_type;          /* setup.basal_stress-type       (lemieux)  */
    int ice_cat_type;               /* thermo.newice_type==4 -> YOUNG_ICE (FE.cpp:1212-1215) */
    int substeps;                   /* dynamics.substeps             (120)      */
    int equal_ridging;              /* age.equal_ridging             (false)    */
    int newice_type;                /* thermo.newice_type            (4)        */
    int use_young_ice_in_myi_reset; /* age.include_young_ice         (true)     */
    int stop_after_substeps;        /* debug: run only the first k sub-cycles (0 = all)           */
    int skip_ow_smoother;           /* debug: skip the open-water smoother (FE.cpp:10578-10611)   */
    int use_coriolis;               /* dynamics.use_coriolis (only zeroes the turning angle, Q3)  */
    double dtime_step;              /* simul.timestep                (200 s)    */
    double ocean_turning_angle_rad; /* dynamics.oceanic_turning_angle (25 deg) * pi/180, FE.cpp:1167-1172 */
    double min_h, min_c;            /* dynamics.min_h (0.05), dynamics.min_c (0.01) */
    double young, nu0, tan_phi;     /* 5.9605e8, 1/3, 0.7 */
    double compr_strength;          /* dynamics.compr_strength (1e10) ALREADY times scale_coef (FE.cpp:6998) */
    double compaction_param;        /* -20 */
    double undamaged_time_relaxation_sigma, exponent_relaxation_sigma;   /* 1e7, 5 */
    double compression_factor, exponent_compression_factor;             /* 10e3, 1.5 */
    double quad_drag_coef_water;    /* 0.0055 */
    double evp_e, evp_Pstar, evp_C, evp_dmin;    /* 2, 27.5e3, 20, 1e-9 */
    double mevp_alpha, mevp_beta;   /* 500, 500 */
    double basal_k1, basal_k2, basal_Cb, basal_u0;   /* 10, 15, 20, 5e-5 */
    /* host-side only (not used by kernels): cohesion recipe, FE.cpp:6995-6999 */
    double C_lab, alea_factor, time_relaxation_damage_days;
} NsxDynParams;

/* One rank's mesh after FiniteElement::distributedMeshProcessing() (FE.cpp:50-143).  Replaces the reads of
 * M_mesh / M_elements / bamgmesh / M_mask_dirichlet / M_neumann_flags inside the path. */
typedef struct NsxMesh {
    int num_nodes;          /* M_num_nodes  (owned + ghost)         FE.cpp:91 */
    int local_ndof;         /* M_local_ndof (owned)                 FE.cpp:87 */
    int num_elements;       /* M_num_elements (owned + ghost)       FE.cpp:84 */
    int local_nelements;    /* M_local_nelements                    FE.cpp:90 */
    const double* coord_x;  /* M_mesh.coordX()  [num_nodes] */
    const double* coord_y;  /* M_mesh.coordY()  [num_nodes] */
    const int* indices;     /* M_mesh.indexTr() [3*num_elements], 1-based local node ids */
    const unsigned char* ghost_nodes;     /* GMSHElement::ghostNodes [3*num_elements]; NULL = derive (id > local_ndof) */
    const unsigned char* mask_dirichlet;  /* M_mask_dirichlet [num_nodes] (owned nodes only, FE.cpp:228-234) */
    const int* neumann_flags;             /* M_neumann_flags, sorted 0-based local node ids */
    int n_neumann_flags;
    const double* nodal_element_connectivity;  /* bamgmesh->NodalElementConnectivity, NaN padded, 1-based */
    int nec_width;                              /* bamgmesh->NodalElementConnectivitySize[1] */
    const double* nodal_connectivity;          /* bamgmesh->NodalConnectivity, last column = count */
    int nc_width;                               /* bamgmesh->NodalConnectivitySize[1] */
    const double* lat;      /* M_mesh.lat() [num_nodes], degrees */
} NsxMesh;

/* Products of FiniteElement::initUpdateGhosts() (FE.cpp:14003-14088) as CSR lists.  NULL for one rank. */
typedef struct NsxHalo {
    int rank, nranks;
    int n_send_peers;       /* M_recipients_proc_id.size() */
    const int* send_peer;   /* M_recipients_proc_id */
    const int* send_ptr;    /* [n_send_peers+1] offsets into send_idx */
    const int* send_idx;    /* M_extract_local_index[peer][:] concatenated (0-based local node ids) */
    int n_recv_peers;       /* M_local_ghosts_proc_id.size() */
    const int* recv_peer;   /* M_local_ghosts_proc_id */
    const int* recv_ptr;
    const int* recv_idx;    /* M_local_ghosts_local_index[peer][:] concatenated */
} NsxHalo;

/* Host arrays named after the FiniteElement members they mirror.  A NULL pointer means "not transferred".
 * nsx_upload copies host -> device, nsx_download device -> host, for every non-NULL member. */
typedef struct NsxFields {
    /* nodal, [2*num_nodes] */
    double* M_VT; double* M_UM; double* M_UT;
    double* M_wind; double* M_ocean;            /* ExternalData::getVector() of M_wind / M_ocean */
    double* M_tau_wi;                           /* optional OASIS wave stress (FE.cpp:10408-10415) */
    double* D_tau_a; double* D_tau_w;           /* outputs */
    /* nodal, [num_nodes] */
    double* M_ssh;
    /* element, [num_elements] */
    double* M_sigma[3]; double* M_damage;
    double* M_conc; double* M_thick; double* M_snow_thick;
    double* M_conc_young; double* M_h_young; double* M_hs_young;
    double* M_thick_myi; double* M_conc_myi; double* M_ridge_ratio;
    double* M_element_depth; double* M_drag_ui; double* M_drag_ui_young;
    double* M_Cohesion; double* M_time_relaxation_damage;
    double* M_surface; double* M_delta_x;       /* outputs of the prep loop (FE.cpp:10239-10240) */
    double* M_shape_coeff;                      /* output, [6*num_elements] element-major like M_shape_coeff[cpt][k] */
    double* D_del_ci_ridge_myi;                 /* output of update() */
    /* outputs of nsx_update_ice_diagnostics (FE.cpp:7860-7900), download only */
    double* D_conc; double* D_thick; double* D_snow_thick;
    double* D_sigma[2];                         /* principal stresses */
    double* D_divergence;
} NsxFields;

/* checkFieldsFast()-style summary computed on the device (FE.cpp:14536-14655). */
typedef struct NsxCheck {
    int n_nan;              /* non-finite entries in VT, sigma, damage, conc, thick */
    int n_speed;            /* owned nodes with |u| > 5 m/s */
    int n_range;            /* elements with damage or conc outside [0,1] or thick < 0 */
    int pad_;
    double max_speed;
} NsxCheck;

/* Local part of FiniteElement::checkRegridding() (FE.cpp:8298-8309): the caller ORs `regrid` over the ranks
 * (boost::mpi::all_reduce of one bool, FE.cpp:8306-8307). */
typedef struct NsxRegrid {
    double min_angle;       /* minAngle(M_mesh, M_UM, 1.) of this rank, degrees (FE.cpp:1795-1806) */
    double min_jacobian;    /* min / max of jacobian(element, M_mesh, M_UM, 1.) over the local triangles */
    double max_jacobian;    /*   (FE.cpp:1824-1839) */
    int flip;               /* (min_jacobian <= 0) && (max_jacobian >= 0) */
    int regrid;             /* (min_angle < regrid_angle) || flip */
} NsxRegrid;

/* ExternalData variables consumed by the path (FE.cpp:10842-10900: M_wind, M_ocean, M_ssh) */
enum { NSX_FORCING_WIND = 0, NSX_FORCING_OCEAN = 1, NSX_FORCING_SSH = 2 };

/* Device times (ms, CUDA events on the handle's stream) of the last nsx_explicit_solve / nsx_update,
 * named after the reference's Timer rows (FE.cpp:10217-10642, 8205-8212). */
typedef struct NsxTiming {
    float prep_ms;          /* "prep elements" + "prep nodes" */
    float subcycle_ms;      /* "sub-time stepping" */
    float ow_smoother_ms;   /* "OW smoother" (incl. tau_w diagnostic) */
    float update_ms;        /* "update" */
    int   n_launches;       /* kernels launched by the last nsx_explicit_solve */
    int   n_substeps;       /* sub-cycles executed */
} NsxTiming;

/* How the sub-cycle loop (FE.cpp:10423-10554) is mapped on the device.  AUTO picks by size:
 *   RESIDENT  one persistent launch per model step, state resident in shared memory / registers (<= ~2.2e5 elements
 *             per GPU), tile-to-tile and GPU-to-GPU synchronisation by release/acquire flags;
 *   DIRECT    element kernel + node kernel per sub-cycle for meshes whose working set fits L2;
 *   TILES     persistent TMA tile pipeline streaming from HBM (large meshes). */
enum { NSX_PATH_AUTO = 0, NSX_PATH_TILES = 1, NSX_PATH_DIRECT = 2, NSX_PATH_RESIDENT = 3 };

/* Creation options (nsx_create uses the defaults).  Nothing in libnsx.so reads environment variables. */
typedef struct NsxCreateOptions {
    int path;           /* NSX_PATH_*                                                     (AUTO) */
    int tile_nodes;     /* owned nodes per tile of the TILES path                         (0 = 208) */
    int max_sms;        /* SMs this handle may occupy; ranks sharing one GPU pass SMs/nranks (0 = all) */
    int use_graph;      /* capture explicitSolve() once into a CUDA graph and replay it   (1) */
    int overlap;        /* multi-GPU TILES / DIRECT: 1 = mailbox exchange inside the sub-cycle kernels (default); 2 = round-1
                         * fused boundary launch with epoch flags overlapped with the interior; 0 = exchange kernel after each sub-cycle */
    int boundary_sms;   /* SMs of that boundary launch                                    (0 = automatic) */
    int ow_skip;        /* multi-GPU TILES / DIRECT smoother: skip exchanges between ranks without open water (1) */
    int pad_;
} NsxCreateOptions;
void nsx_create_options_defaults(NsxCreateOptions* o);

/* ---- life cycle (replaces the member allocation in distributedMeshProcessing / initVariables) ---- */
int nsx_create(const NsxMesh* mesh, const NsxHalo* halo, int device, nsx_handle* out);
int nsx_create_ex(const NsxMesh* mesh, const NsxHalo* halo, int device, const NsxCreateOptions* opt, nsx_handle* out);
int nsx_device_sm_count(int device);        /* multiprocessors of a CUDA device (-1: no such device) */
int nsx_destroy(nsx_handle h);
/* host-only input checks of nsx_create (sizes, index ranges, list ordering); message in nsx_last_error(NULL) */
int nsx_validate_mesh(const NsxMesh* mesh, const NsxHalo* halo);
const char* nsx_last_error(nsx_handle h);      /* h may be NULL: error of the last failed nsx_create */
int nsx_version(void);
int nsx_abi_sizes(int* out, int n);          /* sizeof NsxDynParams, NsxMesh, NsxHalo, NsxFields, NsxCheck, NsxTiming, NsxRegrid, NsxCreateOptions, NsxThermoParams */
/* tile decomposition of the sub-cycle kernel: ntiles, nodes/tile, slots, max local nodes, max slots,
 * boundary tiles, dynamic shared memory bytes, selected path (NSX_PATH_TILES / DIRECT / RESIDENT) */
int nsx_tile_info(nsx_handle h, int* out, int n);
/* host only (no GPU): the tile plan nsx_create would build for this mesh; out[0..9] = ntiles, nodes/tile, slots,
 * max local nodes, max slots, max own slots, max halo slots, max halo nodes, stage bytes, shrink attempts */
int nsx_plan_info(const NsxMesh* mesh, int target_tile_nodes, int wave_ctas, int* out, int n);
/* host only (no GPU): the state-resident plan nsx_create_ex would try on `sms` SMs; out[0..11] = fits, ntiles, nodes/tile,
 * slot space, max slots per tile, max local nodes per tile, shared-memory bytes, export nodes, early own slots, halo
 * slots, own slots, shared-memory limit.  When it does not fit, nsx_last_error(NULL) says why. */
int nsx_resident_plan_info(const NsxMesh* mesh, const NsxHalo* halo, int sms, int* out, int n);
const char* nsx_cfg_last_error(void);        /* message of the last failed nsx_params_from_cfg */

/* ---- options ---- */
void nsx_params_defaults(NsxDynParams* p);                     /* model/options.cpp defaults */
int  nsx_params_from_cfg(const char* path, NsxDynParams* p);    /* nextsim.cfg INI reader, unknown [dynamics] keys rejected */
int  nsx_set_params(nsx_handle h, const NsxDynParams* p);

/* ---- state transfer ---- */
int nsx_upload(nsx_handle h, const NsxFields* f);
int nsx_download(nsx_handle h, NsxFields* f);

/* ---- the path ---- */
int nsx_explicit_solve(nsx_handle h);   /* FiniteElement::explicitSolve()   FE.cpp:10182 */
int nsx_update(nsx_handle h);           /* FiniteElement::update(UM_P)      FE.cpp:3919  */
int nsx_update_ghosts(nsx_handle h);    /* FiniteElement::updateGhosts(M_VT) FE.cpp:13963 */
int nsx_check(nsx_handle h, NsxCheck* out);

/* ---- the callers either side of the path, kept device-resident ----
 * FiniteElement::checkRegridding() without the all_reduce (FE.cpp:8298-8309; minAngle 1795-1816, flip 1824-1839),
 * on the device-resident M_UM: removes the per-step download of M_UM. */
int nsx_check_regridding(nsx_handle h, double regrid_angle, NsxRegrid* out);
/* FiniteElement::updateIceDiagnostics() (FE.cpp:7860-7900): D_conc, D_thick, D_snow_thick, D_sigma[2], D_divergence
 * from the resident state; fetch them with nsx_download.  D_tsurf stays on the host (thermodynamics state). */
int nsx_update_ice_diagnostics(nsx_handle h);
/* ExternalData time interpolation on the device (externaldata.cpp:366-455).  nsx_forcing_load stores
 * Dataset::variables[..].interpolated_data[slot] ([num_nodes] for ssh, [u | v] = [2*num_nodes] for wind / ocean,
 * local numbering); nsx_forcing_apply evaluates ExternalData::getVector() into the resident M_wind / M_ocean / M_ssh:
 *   interp_linear_time: M_factor*(fcoeff[0]*d0[i] + fcoeff[1]*d1[i]) + M_bias_correction, fcoeff from
 *   |current_time - ftime_range[1 or 0]| / |ftime_range[1]-ftime_range[0]| (:368-370, 397-399); else M_factor*d0[i] + bias. */
int nsx_forcing_load(nsx_handle h, int var, int slot, const double* interpolated_data);
int nsx_forcing_apply(nsx_handle h, int var, int interp_linear_time, double current_time, double ftime0, double ftime1,
                      double factor, double bias_correction);
int nsx_synchronize(nsx_handle h);
int nsx_get_timing(nsx_handle h, NsxTiming* out);
void* nsx_get_stream(nsx_handle h);     /* cudaStream_t the handle launches on */

/* ---- FiniteElement::thermo(dt) (FE.cpp:5170-6137) on the device-resident state ----
 * Element-wise: bulk fluxes (OWBulkFluxes 5032-5159, IABulkFluxes 6148-6353 incl. the stability-dependent drag,
 * specificHumidity 4966-5019, albedo 6454-6535), the slab thermodynamics (thermoWinton 6633-6853, thermoIce0 6860-6962),
 * new-ice / lateral-melt redistribution, the slab ocean, melt ponds (6538-6627), healing time, diagnostics and the
 * age / multi-year-ice tracers.  Not covered (the call fails): OASIS coupling (FSD, COUPLED ocean, melt_type 3) and
 * the external AeroBulk library.  Options: model/options.cpp [thermo], [ideal_simul], [age], [dynamics] names. */
typedef struct NsxThermoParams {
    int thermo_type;                 /* setup.thermo-type: 0 zero-layer, 1 winton (enums.hpp:123-127)      (1) */
    int ocean_constant;              /* M_ocean_type == CONSTANT: Qdw / Fdw constants instead of nudging   (1) */
    int Qio_type;                    /* thermo.Qio-type: 0 basic, 1 exchange                               (0) */
    int freezingpoint_type;          /* thermo.freezingpoint-type: 0 linear, 1 unesco                      (0) */
    int newice_type;                 /* thermo.newice_type 1..4                                            (4) */
    int melt_type;                   /* thermo.melt_type 1..2                                              (2) */
    int alb_scheme;                  /* thermo.alb_scheme 1..4                                             (3) */
    int flooding;                    /* thermo.flooding                                                    (1) */
    int use_assim_flux;              /* thermo.use_assim_flux                                              (0) */
    int temp_dep_healing;            /* dynamics.use_temperature_dependent_healing                         (0) */
    int use_meltponds;               /* thermo.use_meltponds                                               (0) */
    int force_neutral_atmosphere;    /* thermo.force_neutral_atmosphere                                    (0) */
    int reset_by_date;               /* age.reset_by_date                                                  (0) */
    int equal_melting;               /* age.equal_melting                                                  (1) */
    int use_young_ice_in_myi_reset;  /* age.include_young_ice                                              (1) */
    int ice_cat_young;               /* M_ice_cat_type == YOUNG_ICE (thermo.newice_type == 4)              (1) */
    /* which forcing variables the atmosphere / ocean datasets provide (ExternalData::isInitialized()) */
    int have_sphuma, have_mixrat, have_Qlw_in, have_snowfr, have_snowfall, have_mld;
    int reset_month, reset_day;      /* age.reset_date "mmdd"                                              (9, 15) */
    double dtime_step;               /* simul.timestep                                                     (200) */
    double ocean_nudge_timeT_days, ocean_nudge_timeS_days;      /* (30, 30) */
    double Qdw_const, Fdw_const;     /* ideal_simul.constant_Qdw / _Fdw                                    (0, 0) */
    double hnull, PhiF, PhiM;        /* thermo.hnull, PhiF, PhiM                                           (0.25, 4, 0.5) */
    double assim_flux_exponent;      /* (1) */
    double constant_mld;             /* ideal_simul.constant_mld                                           (9) */
    double I_0;                      /* thermo.I_0                                                         (0.30) */
    double freeze_days_threshold;    /* age.reset_freeze_days                                              (3) */
    double meltpond_runoff_fraction, meltpond_depth_to_fraction;   /* (0.2, 0.8) */
    double drag_ocean_t, drag_ocean_q;                          /* (0.83e-3, 1.5e-3) */
    double alb_ice, alb_sn, alb_ponds;                          /* (0.538, 0.8256, 0.30) */
    double zref_wind, zref_temp, limiting_lengthscale;          /* (10, 2, 1) */
    double quad_drag_coef_air;       /* dynamics.<atmosphere>_quad_drag_coef_air (FE.cpp:1286-1295)        (0.0049, ASR) */
    double ocean_albedo;             /* thermo.albedoW                                                     (0.07) */
    double ks;                       /* thermo.snow_cond                                                   (0.3096) */
    double freezingpoint_mu;         /* thermo.freezingpoint_mu                                            (0.055) */
    double Csens_io;                 /* thermo.Csens_io                                                    (1e-3) */
    double time_relaxation_damage;   /* dynamics.time_relaxation_damage * 86400 (FE.cpp:1186)              (25 d) */
    double deltaT_relaxation_damage; /* dynamics.deltaT_relaxation_damage                                  (20) */
    double h_young_min, h_young_max; /* thermo.h_young_min, h_young_max                                    (0.05, 0.5) */
} NsxThermoParams;
void nsx_thermo_params_defaults(NsxThermoParams* p);
/* Element fields of thermo() by the reference's member name, host numbering, [num_elements] each:
 *   forcing  M_tair M_mixrat M_dair M_sphuma M_mslp M_Qsw_in M_Qlw_in M_tcc M_precip M_snowfall M_snowfr M_mld
 *            M_ocean_temp M_ocean_salt M_conc_upd
 *   state    M_sst M_sss M_tice0 M_tice1 M_tice2 M_tsurf_young M_del_vi_tend M_freeze_days M_freeze_onset M_conc_summer
 *            M_thick_summer M_fyi_fraction M_age_det M_age M_pond_volume M_lid_volume M_drag_ti M_drag_ti_young
 *            D_pond_fraction    (the ice state itself -- M_conc, M_thick, ... -- goes through NsxFields)
 *   outputs  D_tau_ow D_Qa D_Qsw D_Qlw D_Qsh D_Qlh D_Qo D_Qnosun D_Qsw_ocean D_Qassim D_delS D_fwflux_ice D_fwflux D_brine
 *            D_evap D_rain D_vice_melt D_del_vi_young D_del_hi D_del_hi_young D_newice D_mlt_top D_mlt_bot D_snow2ice
 *            D_albedo D_sialb D_del_ci_mlt_myi D_del_vi_mlt_myi D_del_ci_rplnt_myi D_del_vi_rplnt_myi
 * A field that was never uploaded reads as zero. */
int nsx_thermo_upload(nsx_handle h, const char* name, const double* host);
int nsx_thermo_download(nsx_handle h, const char* name, double* host);      /* any field, the ice state included */
/* n fields in one call: one device arena, one permutation kernel, one stream synchronisation */
int nsx_thermo_upload_many(nsx_handle h, int n, const char* const* names, const double* const* host);
int nsx_thermo_download_many(nsx_handle h, int n, const char* const* names, double* const* host);
/* The forcing members are ExternalData (two time slices, externaldata.cpp:366-455): like nsx_forcing_load / _apply for
 * the nodal forcing, `load` keeps interpolated_data[slot] of one variable on the device and `apply` evaluates
 * M_factor*(fcoeff[0]*d0[i] + fcoeff[1]*d1[i]) + M_bias_correction into the resident member every step. */
int nsx_thermo_forcing_load(nsx_handle h, const char* name, int slot, const double* interpolated_data);
int nsx_thermo_forcing_apply(nsx_handle h, const char* name, int interp_linear_time, double current_time, double ftime0,
                             double ftime1, double factor, double bias_correction);
/* FiniteElement::thermo(dt) at model time `current_time` (decimal days since 1900-01-01, M_current_time) */
int nsx_thermo(nsx_handle h, const NsxThermoParams* p, int dt, double current_time);

/* ---- multi-GPU halo wiring (replaces the Boost.MPI p2p of updateGhosts) ----
 * One process per GPU: every rank exports a 64-byte CUDA IPC handle of its halo window, the host
 * all-gathers them (MPI / torch.distributed, plumbing only) and connects each peer.  Several ranks
 * inside one process (tests, one GPU): nsx_halo_connect_local(). */
#define NSX_IPC_HANDLE_BYTES 64
/* blob for `peer_rank` = [64-byte CUDA IPC handle of my halo window | int num_nodes | int n | n ghost ids
 * (my M_local_ghosts_local_index[peer_rank])]; the peer needs it to store straight into my ghost slots. */
int nsx_halo_blob_size(nsx_handle h, int peer_rank);
int nsx_halo_blob(nsx_handle h, int peer_rank, unsigned char* out);
int nsx_halo_connect_blob(nsx_handle h, int peer_rank, const unsigned char* blob_from_peer);
int nsx_halo_connect_local(nsx_handle h, int peer_rank, nsx_handle peer);
int nsx_halo_finalize(nsx_handle h);
/* lock-step explicitSolve of several ranks that live in this process (one stream-ordered device) */
int nsx_group_explicit_solve(int n, nsx_handle* handles);

/* ---- remesh-time host library (no GPU) ----
 * One rank's view of the partitioned mesh, built without std::map / bimap: the msh-2.2 reader
 * (GmshMesh::readFromFileASCII / readFromFileBinary, core/src/gmshmesh.cpp:133-409, 411-712), GmshMesh::nodalGrid()
 * (core/src/gmshmesh.cpp:856-1498), the lists of FiniteElement::initUpdateGhosts() (FE.cpp:14003-14088, derived from
 * the file instead of MPI exchanges), bcMarkedNodes() (FE.cpp:150-271) and bamg's NodalElementConnectivity /
 * NodalConnectivity (contrib/bamg/src/Mesh.cpp:526-537, 583-629, 798-865).  Numbering is bit-exact with the
 * reference's; nsx_partmesh_views() fills the structs nsx_create() takes. */
typedef struct nsx_partmesh* nsx_partmesh_handle;
/* `path` = <exporter_path>/par<N><mesh.filename> (FE.cpp:1430-1434); format "ascii"|"binary" (mesh.fileformat),
 * ordering "gmsh"|"bamg" (mesh.ordering, swaps vertices 2 and 3) */
int nsx_partmesh_read(const char* path, const char* format, const char* ordering, int rank, int nranks,
                      nsx_partmesh_handle* out);
/* same from memory: root mesh (1-based triangles) + per-element partition (0-based) + ghost partitions (CSR) */
int nsx_partmesh_build(int nn, const double* x, const double* y, int ne, const int* tri, const int* partition,
                       const int* ghost_ptr, const int* ghost_val, int rank, int nranks, nsx_partmesh_handle* out);
/* M_dirichlet_flags_root / M_neumann_flags_root: 1-based root node ids */
int nsx_partmesh_bc_marked_nodes(nsx_partmesh_handle h, const int* dirichlet_flags_root, int n_dirichlet,
                                 const int* neumann_flags_root, int n_neumann);
int nsx_partmesh_set_lat(nsx_partmesh_handle h, const double* lat_local);     /* M_mesh.lat(), local numbering */
/* GmshMesh::lat() / lon() (core/src/gmshmesh.cpp:1800-1824, called by explicitSolve at FE.cpp:10351): inverse polar
 * stereographic map of n points in map units with the projection of `mppfile` (mesh.mppfile: NpsNextsim.mpp /
 * NpsASR.mpp, mapx positional format).  lat or lon may be NULL.  nsx_partmesh_lat_from_mpp fills the handle's lat. */
int nsx_mapx_latlon(const char* mppfile, int n, const double* x, const double* y, double* lat, double* lon);
int nsx_partmesh_lat_from_mpp(nsx_partmesh_handle h, const char* mppfile);
const char* nsx_mapx_last_error(void);
int nsx_partmesh_views(nsx_partmesh_handle h, NsxMesh* mesh, NsxHalo* halo);  /* pointers valid until destroy */
/* ids[0..3]: local node -> root node id (1-based), local node -> reordered global id, local element -> file element
 * number, local element -> partition; sizes[0..3]: global nodes, global triangles, ghost nodes, dirichlet flags */
int nsx_partmesh_ids(nsx_partmesh_handle h, const int** ids, int* sizes);
int nsx_partmesh_destroy(nsx_partmesh_handle h);
const char* nsx_partmesh_last_error(void);

/* ---- pinned host memory for the per-step transfers (cudaHostRegister on the caller's vectors) ---- */
int nsx_host_register(void* p, unsigned long bytes);
int nsx_host_unregister(void* p);

#ifdef __cplusplus
}
#endif
#endif /* NSX_H */
