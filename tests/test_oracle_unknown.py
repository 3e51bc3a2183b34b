"""oracle/unknown.py pinned against OpenCV itself (cv2 is what the reference calls, utils/utils.py:27,53)."""
import cv2
import numpy as np
import pytest

from oracle import unknown as U


@pytest.mark.parametrize("k", range(1, 30))
def test_ellipse_matches_cv2(k):
    assert (cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)) == U.ellipse_kernel(k)).all()


def test_survey_row_spans():
    # row widths dumped from OpenCV 4.13
    w = lambda k: [j2 - j1 for j1, j2 in U.ellipse_spans(k)]
    assert w(15) == [1, 9, 11, 13, 13, 15, 15, 15, 15, 15, 13, 13, 11, 9, 1]
    assert w(13) == [1, 7, 9, 11, 13, 13, 13, 13, 13, 11, 9, 7, 1]
    assert w(7) == [1, 5, 7, 7, 7, 5, 1]


@pytest.mark.parametrize("k", [1, 2, 3, 6, 7, 13, 14, 15, 22, 29])
def test_dilate_matches_cv2(k):
    rng = np.random.RandomState(k)
    for shape, p in (((64, 96), 0.98), ((33, 17), 0.9), ((8, 8), 0.5)):
        u = (rng.rand(*shape) > p).astype(np.uint8)
        ref = cv2.dilate(u, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)))
        assert (ref == U.dilate(u, k)).all()


def test_compute_unknown_thresholds_and_edges():
    a = np.zeros((2, 3, 32, 40), np.float32)
    a[0, 0, 5, 5] = 0.5
    a[0, 1, 0, 0] = 1.0 / 255.0          # not strictly greater -> inactive
    a[0, 2, 31, 39] = 0.99               # active at the corner
    a[1, 0] = 1.0                        # saturated -> inactive
    out = U.compute_unknown(a, [7] * 6)
    assert out.dtype == np.uint8 and out.shape == a.shape
    assert out[0, 0].sum() == U.ellipse_kernel(7).sum()
    assert out[0, 1].sum() == 0 and out[1].sum() == 0
    assert out[0, 2, 31, 39] == 1 and out[0, 2, 28, 39] == 1 and out[0, 2, 27, 39] == 0
    # empty input
    assert U.compute_unknown(np.zeros((0, 4, 4), np.float32), []).shape == (0, 4, 4)


def test_active_sites_order_and_downscale():
    roi = np.zeros((2, 8, 8), np.uint8)
    roi[1, 3, 4] = roi[0, 7, 7] = roi[0, 0, 0] = 1
    s = U.active_sites(roi)
    assert s.tolist() == [[0, 0, 0], [0, 7, 7], [1, 3, 4]]
    s2, shp = U.downscale_sites(s, 8, 8)
    # q covers inputs 2q-1..2q+1: (0,0)->q(0,0); (7,7)->q(3,3) only (q=4 is out of range); (3,4)->qy in{1,2}, qx=2
    assert shp == (4, 4)
    assert s2.tolist() == [[0, 0, 0], [0, 3, 3], [1, 1, 2], [1, 2, 2]]
