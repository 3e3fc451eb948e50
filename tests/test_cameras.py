"""Host-side camera conventions against the golden subset of the reference's rig
(tests/golden/cameras_subset.json, generated from core/dataset/camera_full_calibration.json)."""
import json
import os

import numpy as np

from sigman_release_b200 import cameras

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cameras_subset.json")))


def test_orbit_matches_reference_rig():
    assert GOLD["num_views_in_source"] == cameras.NUM_ORBIT_VIEWS
    for key, cam in GOLD["views"].items():
        w2c = cameras.orbit_w2c(int(key))
        np.testing.assert_allclose(w2c[:3, :3], np.array(cam["R"]), atol=2e-6, err_msg=key)
        np.testing.assert_allclose(w2c[:3, 3], np.array(cam["T"]), atol=2e-6, err_msg=key)
        assert cam["K"][0][0] == 1100.0 and cam["K"][0][2] == 512.0 and cam["height"] == 1024


def test_projection_and_fov():
    P = cameras.sigman_projection()
    assert P[0, 0] == 2.1484375 and P[1, 1] == 2.1484375 and P[3, 2] == 1.0 and P[0, 2] == 0.0
    np.testing.assert_allclose(P[2, 2], 100 / 99.9)
    np.testing.assert_allclose(P[2, 3], -10 / 99.9)
    # tan(FoVy/2) = 512/1100: the rasteriser's tanfov and P[0,0] describe the same pinhole
    np.testing.assert_allclose(cameras.tan_half_fov(), 512 / 1100, rtol=1e-12)
    np.testing.assert_allclose(1.0 / P[0, 0], cameras.tan_half_fov(), rtol=1e-12)


def test_rasterizer_matrix_layout():
    w2c = cameras.orbit_w2c(37)
    vm, pm, cp = cameras.rasterizer_matrices(w2c)
    # flat arrays are column-major W2C and P @ W2C
    p = np.array([0.1, -0.2, 0.3, 1.0])
    flat = vm.reshape(-1)
    got = np.array([flat[0] * p[0] + flat[4] * p[1] + flat[8] * p[2] + flat[12],
                    flat[1] * p[0] + flat[5] * p[1] + flat[9] * p[2] + flat[13],
                    flat[2] * p[0] + flat[6] * p[1] + flat[10] * p[2] + flat[14]])
    np.testing.assert_allclose(got, (w2c @ p)[:3], atol=1e-6)
    fp = pm.reshape(-1)
    w = fp[3] * p[0] + fp[7] * p[1] + fp[11] * p[2] + fp[15]
    np.testing.assert_allclose(w, (w2c @ p)[2], atol=1e-6)          # p_hom.w = view-space z
    np.testing.assert_allclose(np.linalg.norm(cp), cameras.ORBIT_RADIUS, rtol=1e-6)
    V, PM, C = cameras.orbit_cameras([30, 37])
    assert V.shape == (2, 4, 4) and PM.shape == (2, 4, 4) and C.shape == (2, 3)
