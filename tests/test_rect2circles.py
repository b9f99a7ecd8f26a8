"""Circle cover of the body rectangle (reference: confrez/control/rect2circles.py) against hand-computed values."""
import numpy as np

from conflict_rez_b200.control.rect2circles import circle_offsets, circle_pair_rows, v2c, v2c_ca
from conflict_rez_b200.pytypes import VehicleState
from conflict_rez_b200.vehicle_types import VehicleBody


def test_offsets_and_centres_of_the_reference_example():
    body = VehicleBody()
    np.testing.assert_allclose(circle_offsets(body), [0.2, 0.95, 1.7, 2.45], atol=1e-15)
    # the reference's own demo pose (rect2circles.py main): x = 1, y = 2, psi = pi / 4
    st = VehicleState()
    st.x.x, st.x.y, st.e.psi = 1.0, 2.0, np.pi / 4
    c = v2c(st, body)
    assert len(c) == body.num_circles == 4
    h = np.sqrt(0.5)
    for (xc, yc, r), o in zip(c, [0.2, 0.95, 1.7, 2.45]):
        assert abs(xc - (1 + o * h)) < 1e-14 and abs(yc - (2 + o * h)) < 1e-14 and r == 0.9
    # the discs cover the centre line from bumper to bumper up to the corner caps: first / last disc reach past -lr and lf
    assert 0.2 - 0.9 <= -body.lr + 1e-12 and 2.45 + 0.9 >= body.lf - 1e-12


def test_vectorised_centres_and_pair_rows():
    body = VehicleBody()
    rng = np.random.default_rng(0)
    x, y, psi = rng.normal(size=(3, 5)), rng.normal(size=(3, 5)), rng.uniform(-np.pi, np.pi, size=(3, 5))
    xcs, ycs = v2c_ca(x, y, psi, body)
    assert xcs.shape == ycs.shape == (3, 5, 4)
    # centres lie on the heading line through the rear axle at the stated offsets
    np.testing.assert_allclose((xcs - x[..., None]) * np.cos(psi)[..., None] + (ycs - y[..., None]) * np.sin(psi)[..., None],
                               np.broadcast_to(circle_offsets(body), (3, 5, 4)), atol=1e-12)
    np.testing.assert_allclose(-(xcs - x[..., None]) * np.sin(psi)[..., None] + (ycs - y[..., None]) * np.cos(psi)[..., None], 0.0, atol=1e-12)
    # two parallel vehicles d metres apart side by side: matching discs are exactly d apart, the rows' minimum is d^2 - (w + buffer)^2
    for d in (1.5, 2.0, 3.0):
        g = circle_pair_rows(0.0, 0.0, 0.3, -d * np.sin(0.3), d * np.cos(0.3), 0.3, body, d_buffer=0.2)
        assert g.shape == (4, 4)
        np.testing.assert_allclose(np.diag(g), d * d - 4.0, atol=1e-12)
        assert abs(g.min() - (d * d - 4.0)) < 1e-12
        np.testing.assert_allclose(g[0, 3], d * d + 2.25**2 - 4.0, atol=1e-12)
