"""X-ray spectra used by the projector.

``spectrums[name]`` is an ``[n_bins, 2]`` float64 array of (energy [eV], photons / (mAs mm^2)),
the same three spectra the reference ships (reference: deepdrr/projector/spectral_data.py:463;
91 / 151 / 211 bins).  The numbers are stored in ``data/spectra.npz`` (tools/gen_tables.py).
Note the last bin of every spectrum carries a *negative* count; it is kept.
"""
import os

import numpy as np

_z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spectra.npz"))
spectrums = {k: _z[k] for k in _z.files}


def get_spectrum(spectrum):
    """Reference: deepdrr/projector/projector.py:260-279 (``_get_spectrum``)."""
    if isinstance(spectrum, np.ndarray):
        return spectrum
    elif isinstance(spectrum, str):
        if spectrum not in spectrums:
            raise KeyError(f"unrecognized spectrum: {spectrum}")
        return spectrums[spectrum]
    else:
        raise TypeError(f"unrecognized spectrum type: {type(spectrum)}")


def spectrum_tables(spectrum_arr: np.ndarray):
    """(energies [keV] f32, pdf f32) exactly as projector.py:1659-1673 builds them."""
    energies = np.ascontiguousarray(spectrum_arr[:, 0].copy() / 1000, dtype=np.float32)
    pdf = np.ascontiguousarray((spectrum_arr[:, 1] / np.sum(spectrum_arr[:, 1])).copy(), dtype=np.float32)
    return energies, pdf
