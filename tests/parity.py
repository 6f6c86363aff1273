"""Error metrics shared by the parity tests (tolerances are BASELINE.json's north_star: 1e-5 relative RMS and
1e-4 max relative error on per-particle accelerations and potentials; list counts bit-exact)."""
import numpy as np

RMS_TOL = 1e-5
MAX_TOL = 1e-4


def acc_errors(a, a_ref):
    """(relative RMS, max relative) of per-particle |da|/|a_ref|."""
    d = np.linalg.norm(a - a_ref, axis=1)
    nrm = np.linalg.norm(a_ref, axis=1)
    rel = d / nrm
    return float(np.sqrt(np.mean(rel ** 2))), float(rel.max())


def pot_errors(p, p_ref):
    """(relative RMS, max relative) of per-particle potentials, each error measured against max(|phi_i|, rms(phi)):
    potentials change sign in a periodic box (the reference's own v_sqrt1 / native-sqrt build variants already differ
    by 1e-3 relative on particles whose potential is ~0), so the floor is the RMS potential of the case.  In nearly
    uniform periodic boxes the tree part of the potential (27 images, |phi_tree| ~ 21 for M = L = 1) cancels against
    the Ewald term to |phi| ~ 1e-2; the GPU path evaluates the cell monopoles of periodic runs in FP64 (k_eval MONO64)
    precisely so that this strict floor holds there too."""
    floor = np.sqrt(np.mean(p_ref ** 2))
    rel = np.abs(p - p_ref) / np.maximum(np.abs(p_ref), floor)
    return float(np.sqrt(np.mean(rel ** 2))), float(rel.max())
