"""Pins the CPU oracle to everything the reference's own tests assert for the detect path
(SURVEY.md section 4 / 8c).  CPU only."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

from conftest import FIXTURE_NAMES, GOLDEN

REFERENCE_COUNTS = {"iphone": 66, "EuRoC": 36, "TUM_VI": 36, "right": 36, "r45": 36, "top": 36,
                    "two_boards": 72}  # tests/test_detector.rs:26-32


@pytest.mark.parametrize("name", sorted(REFERENCE_COUNTS))
def test_reference_tag_counts(oracle, images, name):
    """tests/test_detector.rs:6-33: detect(img).len() == expected."""
    assert len(oracle.detect(images[name])) == REFERENCE_COUNTS[name]


def test_kornia_path_count(oracle, images):
    """tests/test_detector.rs:35-43: detect_kornia(iphone as u8c3) -> 66 (same bytes as RGB8)."""
    assert len(oracle.detect(images["iphone"])) == 66


@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_oracle_regression_against_committed_golden(oracle, images, expected, name):
    e = expected[name]["oracle"]
    tags = oracle.detect(images[name])
    assert sorted(tags) == e["ids"]
    for k, c in e["corners"].items():
        assert np.array_equal(tags[int(k)], np.array(c, np.float32))
    fe = oracle.front_end(images[name])
    assert int(np.float32(fe["min"]).view(np.uint32)) == e["min_response_bits"]
    assert len(fe["centers"]) == e["n_clusters"]
    assert len(fe["raw"]) == e["n_raw"] and len(fe["refined"]) == e["n_refined"]
    assert hashlib.sha256(fe["blur"].tobytes()).hexdigest() == e["blur_sha256"]
    assert hashlib.sha256(fe["resp"].tobytes()).hexdigest() == e["resp_sha256"]
    assert hashlib.sha256(fe["labels"].tobytes()).hexdigest() == e["labels_sha256"]


# ---- src/math_util.rs:35-90 ---------------------------------------------------------------
def test_find_xy(oracle):
    xy = np.zeros(2, np.float32)
    oracle.lib().orc_find_xy(1.0, 1.0, -2.0, 1.0, -1.0, 0.0, xy.ctypes.data)
    assert abs(xy[0] - 1.0) < 1e-6 and abs(xy[1] - 1.0) < 1e-6


@pytest.mark.parametrize("a,b,want", [(0, 0, 0), (0, 90, 90), (0, 45, 45), (0, 180, 0), (10, 20, 10)])
def test_theta_distance_degree(oracle, a, b, want):
    assert abs(oracle.lib().orc_theta_distance_degree(a, b) - want) < 1e-6


def test_cross_dot_angle(oracle):
    L = oracle.lib()
    assert abs(L.orc_cross(1, 0, 0, 1) - 1.0) < 1e-6 and abs(L.orc_cross(0, 1, 1, 0) + 1.0) < 1e-6
    assert abs(L.orc_dot(1, 0, 0, 1)) < 1e-6 and abs(L.orc_dot(1, 0, 1, 1) - 1.0) < 1e-6
    assert abs(L.orc_angle_degree(1, 0, 0, 1) - 90.0) < 1e-6
    assert abs(L.orc_angle_degree(1, 0, 1, 1) - 45.0) < 1e-6


# ---- src/saddle.rs:75-174 -------------------------------------------------------------------
def test_is_valid_quad(oracle):
    d0, s1, d1 = (10, 0, 0, 0, 0), (10, 10, 0, 0, 0), (0, 10, 0, 0, 0)
    assert not oracle.is_valid_quad((0, 0, 0, 45.0, 0), d0, s1, d1)
    assert oracle.is_valid_quad((0, 0, 0, 135.0, 0), d0, s1, d1)


# ---- src/image_util.rs:238-317 ----------------------------------------------------------------
def test_tag_affine_last_row(oracle):
    h = oracle.tag_affine([(0, 0), (0, 10), (10, 10), (10, 0)], 10, 0.0)
    assert h.shape == (3, 3) and abs(h[2, 0]) < 1e-6 and abs(h[2, 1]) < 1e-6 and abs(h[2, 2] - 1) < 1e-6
    # source corners (0,0),(0,9),(9,9),(9,0) -> scale 10/9, no shear
    assert abs(h[0, 0] - 10 / 9) < 1e-5 and abs(h[1, 1] - 10 / 9) < 1e-5 and abs(h[0, 1]) < 1e-6


def test_hessian_response_impulse(oracle):
    img = np.zeros((5, 5), np.float32)
    img[2, 2] = 10.0
    resp = oracle.hessian_response(img)
    assert resp[2, 2] == 400.0  # lxx = lyy = -20, lxy = 0 (image_util.rs:284-288)
    assert (resp[0] == 0).all() and (resp[:, 0] == 0).all() and (resp[4] == 0).all()


def test_pixel_bfs(oracle):
    img = np.full((5, 5), 100.0, np.float32)
    img[2, 2] = 10.0
    img[3, 2] = 10.0  # put_pixel(2, 3): x = 2, y = 3
    cluster = oracle.pixel_bfs(img, 2, 2, 50.0)
    assert len(cluster) == 2 and (2, 2) in cluster and (2, 3) in cluster
    assert img[2, 2] == np.finfo(np.float32).max


# ---- src/tag_families.rs + detector.rs:124-169 ---------------------------------------------------
def test_codebook_pinned(oracle, pkg):
    with open(os.path.join(GOLDEN, "codebook_sha256.json")) as f:
        sha = json.load(f)
    for name, fam in [("T16H5", pkg.TagFamily.T16H5), ("T25H7", pkg.TagFamily.T25H7),
                      ("T25H9", pkg.TagFamily.T25H9), ("T36H11", pkg.TagFamily.T36H11)]:
        for codes in (oracle.family_info(name.lower())["codes"], [int(c) for c in pkg.family_info(fam)["codes"]]):
            assert len(codes) == sha["counts"][name]
            assert hashlib.sha256(b"".join(int(v).to_bytes(8, "little") for v in codes)).hexdigest() == sha[name]
    assert (oracle.family_info("t36h11b1")["edge"], oracle.family_info("t36h11b1")["border"]) == (6, 1)
    for fam, want in [("t16h5", (4, 2, 1)), ("t25h7", (5, 2, 2)), ("t25h9", (5, 2, 2)), ("t36h11", (6, 2, 3))]:
        i = oracle.family_info(fam)
        assert (i["edge"], i["border"], i["hamming"]) == want  # detector.rs:369-405


def test_rotate_bits_known_answers(oracle):
    seq = [0xD5D628584, 0xC02CDCEA4, 0x21A146BAB, 0x2573B3403]  # SURVEY.md a-12, code 0 of T36H11
    for a, b in zip(seq, seq[1:] + seq[:1]):
        assert oracle.rotate_bits(a, 6) == b


def test_best_tag(oracle):
    codes = oracle.family_info("t36h11")["codes"]
    assert oracle.best_tag(codes[17], 3) == (17, 0)
    assert oracle.best_tag(codes[17] ^ 0b101, 3) == (17, 0)           # 2 bit errors accepted
    assert oracle.best_tag(codes[17] ^ 0b10101, 3) is None or oracle.best_tag(codes[17] ^ 0b10101, 3)[0] != 17
    r = oracle.rotate_bits(codes[5], 6)
    # a tag seen rotated once needs three more rotations to come back
    assert oracle.best_tag(r, 3) == (5, 3)


def test_tag_family_from_str(pkg):
    """src/tag_families.rs:661-685"""
    F = pkg.TagFamily
    assert F.from_str("t36h11") == F.T36H11 and F.from_str("T36H11") == F.T36H11
    assert F.from_str("t16h5") == F.T16H5 and F.from_str("t25h9") == F.T25H9
    assert F.from_str("t36h11b1") == F.T36H11B1
    with pytest.raises(ValueError):
        F.from_str("invalid")


def test_blur_taps_are_the_pinned_constants(oracle):
    """glibc expf taps; the CUDA kernels hard-code these bit patterns."""
    want = [0x3d160c53, 0x3de3e72b, 0x3e5df27d, 0x3e8a96da, 0x3e5df27d, 0x3de3e72b, 0x3d160c53]
    got = [struct.unpack("<I", struct.pack("<f", float(t)))[0] for t in oracle.blur_taps(1.5)]
    assert got == want


def test_blur_matches_naive_definition(oracle):
    rng = np.random.default_rng(0)
    img = rng.random((23, 31), dtype=np.float32)
    k = oracle.blur_taps(1.5)
    tmp = np.zeros_like(img)
    for y in range(img.shape[0]):
        for x in range(img.shape[1]):
            v = np.float32(0)
            for i in range(7):
                v = np.float32(v + np.float32(img[y, min(max(x + i - 3, 0), img.shape[1] - 1)] * k[i]))
            tmp[y, x] = v
    out = np.zeros_like(img)
    for y in range(img.shape[0]):
        for x in range(img.shape[1]):
            v = np.float32(0)
            for i in range(7):
                v = np.float32(v + np.float32(tmp[min(max(y + i - 3, 0), img.shape[0] - 1), x] * k[i]))
            out[y, x] = v
    assert np.array_equal(out, oracle.gaussian_blur(img, 1.5))


def test_labels_are_4_connected_components_in_raster_order(oracle):
    from scipy import ndimage
    rng = np.random.default_rng(1)
    resp = rng.standard_normal((64, 80)).astype(np.float32)
    resp[0, :] = resp[-1, :] = resp[:, 0] = resp[:, -1] = 0.0
    thr = np.float32(resp.min() * 0.3)
    labels, centers, sizes = oracle.clusters(resp, thr)
    ref, n = ndimage.label(resp < thr, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    assert n == len(centers)
    # canonical relabel: order components by their first raster pixel
    first = {}
    for idx, l in enumerate(ref.ravel()):
        if l and l not in first:
            first[l] = len(first)
    canon = np.vectorize(lambda l: first.get(l, -1))(ref)
    assert np.array_equal(canon, labels)
    ys, xs = np.nonzero(labels == 0)
    assert centers[0, 0] == np.float32(xs.sum()) / np.float32(len(xs))


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_point_index_equals_linear_scan(oracle, mode):
    """The oracle's nearest-neighbour index (a uniform bucket grid standing in for kdtree 0.8's
    `nearest`, detector.rs:550, :592-595; board.rs:88, :192-216) returns exactly the linear scan's
    (d2, index) lists: uniform, clustered-with-duplicates and collinear point sets, k = 1 / 3 / 50,
    queries inside and outside the cloud and on the points themselves (exact ties)."""
    for n, k in ((0, 3), (5, 3), (20, 50), (126, 3), (270, 50), (270, 3), (2300, 3), (2300, 50), (12000, 1)):
        assert oracle.selftest_point_index(n, 1500, k, mode, 7 + n) == 0, (mode, n, k)
