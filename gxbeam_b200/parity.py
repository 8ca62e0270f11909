"""Component-wise relative error used by the parity tests, `__graft_entry__.smoke()` and bench.py's CPU leg.

BASELINE.json's north_star asks that "states, reactions and eigenvalues agree to a relative tolerance of 1e-9".  A norm-wise check
(max|Δ| / max|ref| over a whole state vector) leaves small components unconstrained: a rotation of 1e-3 rad next to a metre-scale
displacement could be wrong in its third digit.  Here every component is held to its OWN magnitude,

    err_i = |x_i - ref_i| / max(|ref_i|, FLOOR_FRAC * scale_g),     scale_g = max |ref| over the component's physical group g,

the floor only catching components that are zero (or round-off) in exact arithmetic: below FLOOR_FRAC of its floor class's scale a
value is compared absolutely, to rtol * FLOOR_FRAC * scale_g (1e-14 of the class scale at rtol = 1e-9).  Floor classes are coarse on
purpose -- kinematic quantities (displacements and rotations), velocities, loads (forces and moments): in a pure tip-moment case EVERY
force is zero in exact arithmetic, so a forces-only class would compare round-off with round-off.
"""
from __future__ import annotations

import numpy as np

FLOOR_FRAC = 1e-5


def slot_groups(n, kind):
    """Group id of every entry of a chain's state vector: static [p(6) e(6)]..., dynamic [p(12) e(6)]...."""
    stride = 12 if kind == 0 else 18
    return (np.arange(n) % stride) // 3


def slot_supergroups(n, kind):
    """Floor classes of the slots: 0 = point displacement / rotation slots, 1 = point velocity slots (dynamic), 2 = element resultants."""
    stride = 12 if kind == 0 else 18
    k = np.arange(n) % stride
    return np.where(k < 6, 0, 2) if kind == 0 else np.where(k < 6, 0, np.where(k < 12, 1, 2))


def componentwise_error(x, ref, groups=None, floor_frac=FLOOR_FRAC):
    """max_i |x_i - ref_i| / max(|ref_i|, floor_frac * scale_g(i)): every component against its OWN magnitude, floored at `floor_frac` of the
    scale (max |ref|, per leading index) of its floor class.  x, ref: [..., n]; groups: [n] integer floor-class ids (None = one class)."""
    x = np.asarray(x, float)
    ref = np.asarray(ref, float)
    n = ref.shape[-1]
    xr = x.reshape(-1, n)
    rr = ref.reshape(-1, n)
    if not np.all(np.isfinite(xr)):
        return float("inf")
    g = np.zeros(n, int) if groups is None else np.asarray(groups)
    worst = 0.0
    for gid in np.unique(g):
        idx = g == gid
        a, r = xr[:, idx], rr[:, idx]
        scale = np.abs(r).max(axis=1, keepdims=True)
        denom = np.maximum(np.abs(r), floor_frac * np.maximum(scale, 1e-300))
        worst = max(worst, float((np.abs(a - r) / denom).max()))
    return worst


def state_error(x, ref, kind):
    """Component-wise error of chain state vectors (static kind = 0, dynamic kind = 1)."""
    return componentwise_error(x, ref, slot_supergroups(np.asarray(ref).shape[-1], kind))
