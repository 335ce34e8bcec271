"""Property tests of the oracle's geometric kernels (hypothesis): the closest-point query of two segments
(SimToolbox/Collision/DCPQuery.hpp:199-308 restated) and the boundary projections.  CPU only."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

finite = st.floats(min_value=-5.0, max_value=5.0, allow_nan=False, allow_infinity=False, width=64)
point = st.tuples(finite, finite, finite).map(np.array)


@settings(max_examples=300, deadline=None)
@given(point, point, point, point)
def test_segment_segment_distance_properties(oracle, p0, p1, q0, q1):
    d, P, Q, s, t = oracle.dcp_segseg(p0, p1, q0, q1)
    # the reported points lie on the segments at the reported parameters and realise the reported distance
    assert 0.0 <= s <= 1.0 and 0.0 <= t <= 1.0
    np.testing.assert_allclose(P, p0 + s * (p1 - p0), atol=1e-12)
    np.testing.assert_allclose(Q, q0 + t * (q1 - q0), atol=1e-12)
    assert abs(np.linalg.norm(P - Q) - d) < 1e-12
    # symmetric in the two segments, invariant under reversal of a segment and under translation
    d2 = oracle.dcp_segseg(q0, q1, p0, p1)[0]
    d3 = oracle.dcp_segseg(p1, p0, q0, q1)[0]
    sh = np.array([0.25, -1.5, 3.0])
    d4 = oracle.dcp_segseg(p0 + sh, p1 + sh, q0 + sh, q1 + sh)[0]
    scale = 1e-9 * (1 + d)
    assert abs(d - d2) < scale and abs(d - d3) < scale and abs(d - d4) < scale
    # no sampled pair of points is closer than the reported minimum
    u = np.linspace(0, 1, 9)
    A = p0[None] + u[:, None] * (p1 - p0)[None]
    B = q0[None] + u[:, None] * (q1 - q0)[None]
    dm = np.sqrt(((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)).min()
    assert d <= dm + 1e-9


@settings(max_examples=200, deadline=None)
@given(point, st.sampled_from(["sphere", "wall", "tube"]), st.booleans())
def test_boundary_projection_is_idempotent(oracle, q, kind, inside):
    b = oracle.make_boundaries([dict(type=kind, center=[0.5, -0.25, 1.0], axis=[1.0, 2.0, -1.0], radius=1.7, inside=inside)])[0]
    ctr = np.array([0.5, -0.25, 1.0])
    ax = b["axis"]
    if kind == "sphere" and np.linalg.norm(q - ctr) < 1e-3:
        return  # the centre has no projection (the reference divides by the distance to the centre)
    if kind == "tube" and np.linalg.norm((q - ctr) - ((q - ctr) @ ax) * ax) < 1e-3:
        return
    proj, delta = oracle.boundary_project(b, q)
    proj2, delta2 = oracle.boundary_project(b, proj)
    np.testing.assert_allclose(proj2, proj, atol=1e-9)  # a point of the surface projects onto itself
    assert np.linalg.norm(delta2) < 1e-9
