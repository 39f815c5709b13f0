"""Host tables vs golden vectors generated from the reference's own Material / spectral_data
(tools/gen_tables.py, run against /root/reference)."""
import numpy as np
import pytest

from cases import GOLDEN
from deepdrr_b200.material import Material, absorb_coef_table
from deepdrr_b200.spectral_data import get_spectrum, spectrum_tables, spectrums


@pytest.fixture(scope="module")
def gold():
    return np.load(f"{GOLDEN}/absorb_tables.npz")


@pytest.mark.parametrize("name,bins", [("60KV_AL35", 91), ("90KV_AL40", 151), ("120KV_AL43", 211)])
def test_spectrum_and_mu_tables_bit_exact(gold, name, bins):
    mats = [str(m) for m in gold["materials"]]
    energies, pdf = spectrum_tables(spectrums[name])
    assert energies.shape == (bins,) and energies.dtype == np.float32
    assert np.array_equal(energies, gold[name + "::energies"])
    assert np.array_equal(pdf, gold[name + "::pdf"])
    assert pdf[-1] < 0  # the last bin carries a negative count, kept
    table = absorb_coef_table(mats, energies)
    assert table.dtype == np.float32 and table.shape == (bins * len(mats),)
    assert np.array_equal(table, gold[name + "::table"])


def test_compound_string_material(gold):
    m = Material.from_string(str(gold["compound::name"]), compound_string=True)
    assert m.get_coefficients(60.0).mu_over_rho == float(gold["compound::mu60"])
    assert m.get_coefficients(33.3).mu_over_rho == float(gold["compound::mu33"])


def test_material_lookup_and_errors():
    assert abs(Material.from_string("bone").get_coefficients(60.0).mu_over_rho - 0.3148) < 1e-9
    assert Material.from_string("soft tissue").name == "tissue_soft"
    assert Material.from_string("iron").name == "26_Fe_Iron"
    assert Material.from_string("Ti").name == "22_Ti_Titanium"
    with pytest.raises(AttributeError):
        Material.from_string("unobtainium")
    with pytest.raises(ValueError):
        Material.from_string("H0.5O0.1", compound_string=True)


def test_get_spectrum_errors():
    with pytest.raises(KeyError):
        get_spectrum("77KV")
    with pytest.raises(TypeError):
        get_spectrum(3)
    arr = np.array([[20000.0, 1.0], [30000.0, 2.0]])
    assert get_spectrum(arr) is arr
