"""Option branches gcov showed no other parity test reaches (see tests/test_ref_fe_cpu.py, last three cases): the compression
cap of the damage criterion (FE.cpp:4218-4221), no basal stress, the classic ice category in update()."""
import pytest

from nextsim_b200 import cases
from test_gpu_parity import run_both

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["tiles", "direct", "resident"])
def solver_path(request, monkeypatch):
    monkeypatch.setenv("NSX_PATH", request.param)
    return request.param


@pytest.mark.parametrize("over,open_east,young", [
    ({"compr_strength": 2.0e4}, True, True),
    ({"basal_stress_type": 0}, True, True),
    ({"newice_type": 1, "ice_cat_type": 0}, False, False),
])
def test_option_branches(over, open_east, young):
    c = cases.make_case("10km_stable", nranks=1, dyn="bbm", nx=40, open_east=open_east, young=young, substeps=30)
    for k, v in over.items():
        setattr(c.params, k, v)
    run_both(c)
