"""Shared pieces of the thermo() tests: option sets that reach every branch, model times that hit the date logic, and
drivers of the two CPU sides (oracle/ref_fe = the reference's own bodies; oracle.thermo = the host build of the element
function the kernel is compiled from)."""
import datetime

import numpy as np

from nextsim_b200 import synthetic as syn
from oracle import thermo as oth

ALL_FIELDS = syn.THERMO_FORCING + syn.THERMO_ICE + syn.THERMO_STATE + syn.THERMO_DIAG
OUT_FIELDS = syn.THERMO_ICE + syn.THERMO_STATE + syn.THERMO_DIAG


def datenum(y, m, d, frac=0.0):
    """nextsim time: decimal days since 1900-01-01 00:00 (core/include/date.hpp)."""
    return float((datetime.date(y, m, d) - datetime.date(1900, 1, 1)).days) + frac


# name -> (option overrides, current_time, dt, season)
# dt = 200 s: 432 steps per day; step_in_day = 1 + round(432 * frac)
LAST_STEP = 431.0 / 432.0
OPTION_SETS = {
    "defaults": ({}, datenum(2018, 2, 3, 0.25), 200, "mixed"),
    "defaults_summer": ({}, datenum(2018, 7, 3, 0.5), 200, "summer"),
    "zero_layer": (dict(thermo_type=0), datenum(2018, 2, 3, 0.25), 200, "mixed"),
    "newice1_melt1": (dict(newice_type=1, ice_cat_young=0, melt_type=1), datenum(2018, 5, 3, 0.125), 200, "mixed"),
    "newice2_alb1": (dict(newice_type=2, ice_cat_young=0, alb_scheme=1), datenum(2018, 11, 3, 0.75), 200, "winter"),
    "newice3_alb2_ice0": (dict(newice_type=3, ice_cat_young=0, alb_scheme=2, thermo_type=0), datenum(2018, 6, 3, 0.75), 200, "mixed"),
    "alb4_ponds": (dict(alb_scheme=4, use_meltponds=1), datenum(2018, 7, 10, 0.5), 200, "summer"),
    "ponds_alb3": (dict(use_meltponds=1), datenum(2018, 7, 10, 0.5), 200, "mixed"),
    "nudged_ocean": (dict(ocean_constant=0, have_mld=1, Qio_type=1, freezingpoint_type=1), datenum(2018, 3, 3, 0.3), 200, "mixed"),
    "forcing_variants_a": (dict(have_sphuma=1, have_Qlw_in=1, have_snowfr=1), datenum(2018, 3, 3, 0.3), 200, "mixed"),
    "forcing_variants_b": (dict(have_mixrat=1, have_snowfall=1, force_neutral_atmosphere=1), datenum(2018, 3, 3, 0.3), 200, "mixed"),
    "assim_healing": (dict(use_assim_flux=1, assim_flux_exponent=2.0, temp_dep_healing=1), datenum(2018, 1, 3, 0.3), 200, "winter"),
    "healing_ice0_noflood": (dict(temp_dep_healing=1, thermo_type=0, flooding=0), datenum(2018, 1, 3, 0.3), 200, "mixed"),
    "first_step_of_day": ({}, datenum(2018, 2, 3, 0.0), 200, "winter"),
    "last_step_of_day": ({}, datenum(2018, 2, 3, LAST_STEP), 200, "mixed"),
    "last_step_no_young_reset": (dict(use_young_ice_in_myi_reset=0, equal_melting=0), datenum(2018, 8, 20, LAST_STEP), 200, "summer"),
    "sept15_midnight": ({}, datenum(2018, 9, 15, 0.0), 200, "mixed"),
    "aug01_midnight": ({}, datenum(2018, 8, 1, 0.0), 200, "summer"),
    "reset_by_date_hit": (dict(reset_by_date=1, reset_month=10, reset_day=2), datenum(2018, 10, 2, 0.0), 200, "mixed"),
    "reset_by_date_miss": (dict(reset_by_date=1), datenum(2018, 10, 2, 0.5), 200, "mixed"),
    "reset_by_date_last_step": (dict(reset_by_date=1), datenum(2018, 7, 20, LAST_STEP), 200, "summer"),
    "reset_by_date_aug01": (dict(reset_by_date=1), datenum(2018, 8, 1, 0.0), 200, "summer"),
    "dt900": (dict(dtime_step=900.0), datenum(2020, 2, 29, 0.5), 900, "mixed"),
    "classic_winton_newice1": (dict(newice_type=1, ice_cat_young=0, hnull=0.4, PhiM=0.3), datenum(2018, 4, 3, 0.6), 200, "mixed"),
}


def mesh(nx=24):
    gm = syn.make_mesh(nx, 10e3)
    return gm


def make_inputs(name, nx=24, seed=syn.SEED, films=True):
    over, t, dt, season = OPTION_SETS[name]
    p = oth.default_params(**over)
    gm = mesh(nx)
    S = syn.make_thermo_state(gm.ne, gm.nn, seed=seed, young=bool(p.ice_cat_young), season=season, films=films)
    return p, t, dt, gm, S


def run_oracle(p, t, dt, gm, S, steps=1):
    """thermo() `steps` times (model time advancing by dt), returns the fields after the last call."""
    F = {k: S[k] for k in ALL_FIELDS if k in S}
    for s in range(steps):
        F = oth.thermo(p, dt, t + s * dt / 86400.0, gm.tri - 1, gm.nn, S["M_wind"], S["M_VT"], S["M_ocean"], F)
    return F


def run_reference(p, t, dt, gm, S, steps=1):
    """The reference's own thermo() (oracle/ref_fe)."""
    from oracle import ref_fe
    R = ref_fe.RefFE(1)
    R.set_mesh(0, gm.x, gm.y, gm.nn, gm.tri, np.zeros(gm.nn, np.uint8), np.zeros(0, np.int32))
    R.thermo_setup(p, t)
    three = p.thermo_type == 1
    for k in ALL_FIELDS + ("M_wind", "M_VT", "M_ocean"):
        if k in S and (three or k not in ("M_tice1", "M_tice2")):
            R.set(0, k, S[k])
    for s in range(steps):
        R.thermo_setup(p, t + s * dt / 86400.0)
        R.thermo(dt)
    out = {}
    for k in ALL_FIELDS:
        if three or k not in ("M_tice1", "M_tice2"):
            out[k] = R.get(0, k)
    return out
