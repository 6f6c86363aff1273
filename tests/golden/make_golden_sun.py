"""Golden vectors for bDoSun and bComove (pkd.c:3003-3041: the indirect acceleration at the origin, a dummy sink of softening dSunSoft)
from the COMPILED REFERENCE.  Run where /root/reference exists:  python tests/golden/make_golden_sun.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gasoline_b200 import ics  # noqa: E402
from oracle import reflib  # noqa: E402


def case(name):
    """(particles, theta, dSunSoft): the origin inside the cloud, well outside it, and on top of a particle"""
    if name == "inside":
        return ics.plummer(3000, seed=41), 0.7, 0.01
    if name == "offset":
        p = ics.plummer(4000, seed=42)
        return ics.Particles(p.x + 3.0, p.y - 1.0, p.z + 0.5, p.m, p.h, p.period, "plummer_offset"), 0.5, 0.05
    p = ics.plummer(2500, seed=43)
    x, y, z = p.x.copy(), p.y.copy(), p.z.copy()
    x -= x[7]; y -= y[7]; z -= z[7]  # particle 7 sits exactly at the origin: the softened pair term
    return ics.Particles(x, y, z, p.m, p.h, p.period, "plummer_on_particle"), 0.7, 0.02


NAMES = ("inside", "offset", "on_particle")

if __name__ == "__main__":
    out = {}
    for name in NAMES:
        p, theta, soft = case(name)
        r = reflib.RefGravity(p)
        r.build_tree(8, theta, 4)
        a, c3 = r.gravity_sun(soft)
        plain = r.gravity(0, 0, 4, 0, 4)
        r.close()
        out[name + "_aSun"], out[name + "_counts"], out[name + "_sums"] = a, c3, np.array(
            [plain["nActive"], plain["dPartSum"], plain["dCellSum"], plain["dSoftSum"], plain["dFlop"]])
        print(name, a, c3)
    # bComove (pkd.c:2967-2991) on the first case, partially active
    p, theta, _ = case("inside")
    active = (np.arange(p.n) % 3 != 0).astype(np.int32)
    r = reflib.RefGravity(p, active=active)
    r.build_tree(8, theta, 4)
    out["comove_acc"], out["comove_pot"] = r.gravity_comove(0.37)
    out["comove_active"] = r.tree()["active"]
    r.close()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sun.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)
