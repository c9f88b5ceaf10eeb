"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: -m gpu.

Bars (BASELINE.json north_star): blurred image, response, threshold mask, component labels and
tag ids bit-exact; refined corner coordinates within 1e-3 px (they come out bit-identical too;
the tolerance is written where it is used)."""
import os

import numpy as np
import pytest

import synth
from conftest import FIXTURE_NAMES

pytestmark = pytest.mark.gpu
CORNER_TOL = 1e-3  # px, north_star


def assert_tags_match(got, want):
    assert sorted(got) == sorted(want), (sorted(got), sorted(want))
    for k in want:
        assert np.abs(got[k] - want[k]).max() <= CORNER_TOL, (k, got[k], want[k])


def assert_saddles_match(g, o):
    """g: structured array from the GPU, o: N x 5 oracle array."""
    assert len(g) == len(o)
    if len(o) == 0:
        return
    # positions and k use +,-,*,/,sqrt only: bit-identical.  theta/phi pass through acos/atan2: the
    # kernels evaluate glibc's acosf / atan2f operation for operation (csrc/ag_libm.h), the oracle
    # calls the platform's libm as a Rust binary would: the same bits, so every gate downstream
    # (phi in [30, 60], theta differences, round(theta)) sees identical inputs.
    assert np.array_equal(g["x"], o[:, 0]) and np.array_equal(g["y"], o[:, 1])
    assert np.array_equal(g["k"], o[:, 2])
    assert np.array_equal(g["theta"].view(np.uint32), np.ascontiguousarray(o[:, 3]).view(np.uint32))
    assert np.array_equal(g["phi"].view(np.uint32), np.ascontiguousarray(o[:, 4]).view(np.uint32))


def check_stages(det, oracle, img, check_board=True):
    g = det.stages(img)
    o = oracle.front_end(img)
    assert np.array_equal(g["blur"].view(np.uint32), o["blur"].view(np.uint32)), "blur not bit-exact"
    assert np.array_equal(g["resp"].view(np.uint32), o["resp"].view(np.uint32)), "response not bit-exact"
    assert np.float32(g["min"]) == np.float32(o["min"]) and np.float32(g["thr"]) == np.float32(o["thr"])
    assert np.array_equal(g["mask"].astype(bool), o["resp"] < np.float32(o["thr"]))
    assert np.array_equal(g["labels"], o["labels"]), "labels differ"
    if len(o["centers"]) > 16384 and len(g["centers"]) == 16384:
        # more clusters than the default max_clusters: the first 16384 (raster order) are kept and
        # the frame is flagged; nothing downstream is comparable
        assert np.array_equal(g["centers"].view(np.uint32), o["centers"][:16384].view(np.uint32))
        return g, o
    assert len(g["centers"]) == len(o["centers"])
    differ = np.nonzero((g["centers"].view(np.uint32) != o["centers"].view(np.uint32)).any(axis=1))[0]
    if len(differ):
        # Documented deviation: the reference sums pixel coordinates in f32 in its
        # flood-fill order; a component whose coordinate sum reaches 2^24 is no longer exact in
        # f32 and order-dependent.  The CUDA path sums integers.  Only such giant components
        # (never a saddle) may differ, and only in the last digits.
        sizes = np.bincount(o["labels"][o["labels"] >= 0])
        for c in differ:
            assert sizes[c] * max(img.shape[:2]) >= 2 ** 24, (c, sizes[c])
            assert np.allclose(g["centers"][c], o["centers"][c], rtol=1e-4)
        return g, o
    assert_saddles_match(g["raw"], o["raw"])
    assert_saddles_match(g["refined"], o["refined"])
    if check_board:
        oq = oracle.try_find_best_board(o["refined"])
        if oq is None:
            assert len(g["quads"]) == 0
        else:
            assert np.array_equal(g["quads"], oq)
        assert_tags_match(g["tags"], oracle.detect(img))
    return g, o


def test_unorm_conversion_exhaustive(detector):
    o8, o16, r8, r16 = detector._unorm_tables()
    assert np.array_equal(o8.view(np.uint32), r8.view(np.uint32))
    assert np.array_equal(o16.view(np.uint32), r16.view(np.uint32))
    assert np.array_equal(r8, (np.arange(256, dtype=np.float32) / np.float32(255)))
    assert np.array_equal(r16, (np.arange(65536, dtype=np.float32) / np.float32(65535)))


@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_fixture_stage_parity(detector, oracle, images, name):
    check_stages(detector, oracle, images[name])


@pytest.mark.parametrize("name,count", [("iphone", 66), ("EuRoC", 36), ("TUM_VI", 36), ("right", 36),
                                        ("r45", 36), ("top", 36), ("two_boards", 72)])
def test_reference_counts_through_detect(detector, oracle, images, name, count):
    """tests/test_detector.rs:26-32 run against the CUDA path."""
    tags = detector.detect(images[name])
    assert len(tags) == count
    assert_tags_match(tags, oracle.detect(images[name]))


def test_detect_kornia(detector, oracle, images):
    """tests/test_detector.rs:35-43"""
    img = images["iphone"]
    assert len(detector.detect_kornia(img)) == 66
    gray = oracle.to_luma_u8(images["EuRoC"])[:, :, None]
    assert_tags_match(detector.detect_kornia(gray), oracle.detect(images["EuRoC"]))
    with pytest.raises(ValueError):
        detector.detect_kornia(np.zeros((8, 8, 4), np.uint8))


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 3), (5, 7), (8, 9), (9, 33), (31, 65), (40, 130)])
def test_tiny_and_odd_shapes(detector, oracle, shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    check_stages(detector, oracle, img)


@pytest.mark.parametrize("shape", [(5, 8), (3, 12), (17, 120), (9, 124), (130, 244), (33, 1280), (300, 364)])
def test_streaming_dense_kernel_shapes(detector, oracle, shape):
    """Widths that are multiples of 4 take the register-marching K1; strip / chunk edge cases."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    g, o = check_stages(detector, oracle, img, check_board=False)
    # the generic tile kernel must give the same bits
    detector.set_option("dense_variant", 1)
    try:
        g2 = detector.stages(img)
    finally:
        detector.set_option("dense_variant", 0)
    assert np.array_equal(g["blur"].view(np.uint32), g2["blur"].view(np.uint32))
    assert np.array_equal(g["resp"].view(np.uint32), g2["resp"].view(np.uint32))
    assert g["min"] == g2["min"]


@pytest.mark.parametrize("fmt", ["l16", "rgb8"])
@pytest.mark.parametrize("shape", [(5, 8), (17, 120), (9, 124), (130, 244), (140, 364)])
def test_streaming_dense_kernel_formats(detector, oracle, fmt, shape):
    """16-bit gray and RGB8 frames whose width is a multiple of 4 take the register-marching K1 too
    (a lane's 4 pixels are 2 / 3 words): bit-exact with the oracle and with the generic tile kernel."""
    rng = np.random.default_rng(shape[0] * 77 + shape[1] + len(fmt))
    if fmt == "l16":
        img = rng.integers(0, 65536, shape, dtype=np.uint16)
    else:
        img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    g, o = check_stages(detector, oracle, img, check_board=False)
    detector.set_option("dense_variant", 1)
    try:
        g2 = detector.stages(img)
    finally:
        detector.set_option("dense_variant", 0)
    assert np.array_equal(g["blur"].view(np.uint32), g2["blur"].view(np.uint32))
    assert np.array_equal(g["resp"].view(np.uint32), g2["resp"].view(np.uint32))
    assert g["min"] == g2["min"]


@pytest.mark.parametrize("rows", [4, 58, 124, 250, 508])
@pytest.mark.parametrize("shape", [(5, 8), (17, 120), (130, 244), (300, 364), (1024, 1280)])
def test_streaming_dense_kernel_chunk_heights(detector, oracle, shape, rows):
    """The streaming K1 marches down chunks of 6k + 4 rows (the launcher picks the height from the
    batch size); every height gives the bits of the generic tile kernel, for row counts with every
    remainder and chunks shorter than one loop trip."""
    rng = np.random.default_rng(shape[0] + 3 * shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    detector.set_option("dense_variant", 1)
    try:
        g = detector.stages(img)
    finally:
        detector.set_option("dense_variant", 0)
    detector.set_option("k1_chunk_rows", rows)
    try:
        g2 = detector.stages(img)
    finally:
        detector.set_option("k1_chunk_rows", 0)
    assert np.array_equal(g["blur"].view(np.uint32), g2["blur"].view(np.uint32))
    assert np.array_equal(g["resp"].view(np.uint32), g2["resp"].view(np.uint32))
    assert g["min"] == g2["min"] and np.array_equal(g["mask"], g2["mask"])
    if shape[0] <= 300 and rows == 58:
        o = oracle.front_end(img, want_labels=False)
        assert np.array_equal(g2["blur"].view(np.uint32), o["blur"].view(np.uint32))


def test_constant_image_gives_empty_map(detector, oracle):
    for v in (0, 128, 255):
        img = np.full((48, 64), v, np.uint8)
        g, o = check_stages(detector, oracle, img)
        assert g["tags"] == {} and (g["mask"] == 0).all()  # min = 0 -> nothing below threshold
        assert detector.detect(img) == {}


@pytest.mark.parametrize("fmt", ["l8", "l16", "rgb8"])
def test_random_noise_dense_and_labels(detector, oracle, fmt):
    """iid noise: worst case for the labeller (large ragged components)."""
    rng = np.random.default_rng(11)
    if fmt == "l8":
        img = rng.integers(0, 256, (200, 333), dtype=np.uint8)
    elif fmt == "l16":
        img = rng.integers(0, 65536, (200, 333), dtype=np.uint16)
    else:
        img = rng.integers(0, 256, (200, 333, 3), dtype=np.uint8)
    check_stages(detector, oracle, img, check_board=False)


def test_row_stride(detector, oracle, images):
    """Padded rows: the image is a view into a wider buffer."""
    base = images["EuRoC"]
    h, w = base.shape
    buf = np.zeros((h, w + 40), np.uint8)
    buf[:, :w] = base
    view = buf[:, :w]
    assert not view.flags.c_contiguous
    assert_tags_match(detector.detect(view), oracle.detect(base))


@pytest.mark.parametrize("kw", [dict(), dict(dtype=np.uint16), dict(rgb=True)])
def test_synthetic_formats(detector, oracle, kw):
    img = synth.render_board_numpy(640, 480, seed=5, tag_px=44.0, **kw)
    g, o = check_stages(detector, oracle, img)
    assert sorted(g["tags"]) == list(range(36))


def test_empty_and_single_frame_batches(detector, pkg, oracle):
    """n_frames = 0 is a valid call that touches nothing (host and device entry points); a batch
    of one equals detect(); null pointers and negative counts are refused."""
    import torch
    frames = synth.fixture_like_frames(1, 640, 480, seed=31, tag_px=42.0)
    want = oracle.detect(frames[0])
    assert len(want) > 0
    assert detector.detect_batch(frames[:0]) == []
    L = pkg.lib()
    tags = np.zeros((1, 64), pkg.TAG_DTYPE)
    cnt = np.full(1, -7, np.int32)
    vp = lambda a: a.ctypes.data_as(__import__("ctypes").c_void_p)
    rc = L.ag_detect_batch(detector._h, vp(frames), frames.strides[0], 0, 640, 480, 640, pkg.FMT_L8, vp(tags), 64,
                           vp(cnt), None)
    assert rc == pkg.AG_OK and cnt[0] == -7 and not tags["id"].any()
    rc = L.ag_detect_batch(detector._h, vp(frames), frames.strides[0], -1, 640, 480, 640, pkg.FMT_L8, vp(tags), 64,
                           vp(cnt), None)
    assert rc == pkg.AG_ERR_INVALID
    rc = L.ag_detect_batch(detector._h, None, frames.strides[0], 1, 640, 480, 640, pkg.FMT_L8, vp(tags), 64,
                           vp(cnt), None)
    assert rc == pkg.AG_ERR_INVALID
    d_frames = torch.from_numpy(frames).cuda()
    d_tags = torch.zeros((1, 64 * 9), dtype=torch.int32, device="cuda")
    d_cnt = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    detector.detect_batch_device(d_frames.data_ptr(), 0, 640, 480, pkg.FMT_L8, d_tags.data_ptr(), 64, d_cnt.data_ptr())
    torch.cuda.synchronize()
    assert int(d_cnt[0]) == -7
    got = detector.detect_batch(frames)
    assert len(got) == 1
    assert_tags_match(got[0], want)
    assert_tags_match(detector.detect(frames[0]), want)
    detector.detect_batch_device(d_frames.data_ptr(), 1, 640, 480, pkg.FMT_L8, d_tags.data_ptr(), 64, d_cnt.data_ptr())
    torch.cuda.synchronize()
    assert int(d_cnt[0]) == len(want)


def test_detect_batch_matches_oracle_per_frame(detector, oracle):
    frames = synth.fixture_like_frames(7, 640, 480, seed=20, tag_px=42.0)
    frames[3] = 77  # one empty frame in the middle of the batch
    detector.set_option("chunk_frames", 3)  # ragged chunks: 3 + 3 + 1, both pipeline slots used
    try:
        got, status = detector.detect_batch(frames, return_status=True)
    finally:
        detector.set_option("chunk_frames", 32)
    assert (status == 0).all()
    want = oracle.detect_batch(frames)
    assert len(got) == 7 and got[3] == {}
    for g, w in zip(got, want):
        assert_tags_match(g, w)
    assert detector.detect_batch(frames[:0]) == []


def test_capacity_error_is_reported(detector, pkg, images):
    with pytest.raises(RuntimeError):
        detector.detect(images["EuRoC"], cap=5)


def test_operators_blur_and_hessian(detector, oracle):
    """image_util::gaussian_blur_f32 / hessian_response as standalone f32 -> f32 operators."""
    rng = np.random.default_rng(2)
    for shape in [(5, 5), (37, 61), (480, 752)]:
        img = rng.random(shape, dtype=np.float32)
        for sigma in (1.5, 0.8, 2.6):
            assert np.array_equal(detector.gaussian_blur_f32(img, sigma).view(np.uint32),
                                  oracle.gaussian_blur(img, sigma).view(np.uint32))
        assert np.array_equal(detector.hessian_response(img).view(np.uint32),
                              oracle.hessian_response(img).view(np.uint32))
    imp = np.zeros((5, 5), np.float32)
    imp[2, 2] = 10.0
    assert detector.hessian_response(imp)[2, 2] == 400.0  # image_util.rs:284-288


def test_blur_f32_streaming_and_batched(detector, oracle):
    """gaussian_blur_f32 through its streaming kernel (radius 3, width % 4 == 0): the reference's
    bits for every edge case of the strip / chunk geometry, for inputs with negative values and
    signed zeros (the accumulations start from +0.0), and for a device-resident batch."""
    import torch
    rng = np.random.default_rng(11)
    for shape in [(5, 8), (6, 12), (61, 120), (67, 124), (130, 244), (127, 364), (480, 752), (1024, 1280)]:
        img = (rng.random(shape, dtype=np.float32) - 0.5).astype(np.float32)
        img[rng.random(shape) < 0.05] = -0.0
        img[rng.random(shape) < 0.05] = 0.0
        for sigma in (1.5, 1.2):  # both radius 3
            got = detector.gaussian_blur_f32(img, sigma)
            assert np.array_equal(got.view(np.uint32), oracle.gaussian_blur(img, sigma).view(np.uint32)), (shape, sigma)
    n, h, w = 5, 130, 244
    batch = (rng.random((n, h, w), dtype=np.float32) * 3.0 - 1.0).astype(np.float32)
    d_in = torch.from_numpy(batch).cuda()
    guard = 4096  # floats of canary on either side of the output: nothing may be written there
    for sigma in (1.5, 2.6):  # streaming kernel / general two-pass path
        buf = torch.full((guard + n * h * w + guard,), 12345.0, dtype=torch.float32, device="cuda")
        d_out = buf[guard:guard + n * h * w]
        detector.gaussian_blur_f32_device(d_in.data_ptr(), n, w, h, sigma, d_out.data_ptr(),
                                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert bool((buf[:guard] == 12345.0).all()) and bool((buf[guard + n * h * w:] == 12345.0).all())
        out = d_out.cpu().numpy().reshape(n, h, w)
        for i in range(n):
            assert np.array_equal(out[i].view(np.uint32), oracle.gaussian_blur(batch[i], sigma).view(np.uint32)), (i, sigma)
    detector.gaussian_blur_f32_device(0, 0, w, h, 1.5, 0)  # empty batch


def test_refined_saddle_points_api(detector, oracle, images):
    g = detector.refined_saddle_points(images["TUM_VI"])
    assert_saddles_match(g, oracle.front_end(images["TUM_VI"], want_labels=False)["refined"])


def test_other_families(pkg, oracle):
    """T16H5 / T25H7 / T25H9 / T36H11B1: parameters + tables only (detector.rs:369-405)."""
    for fam, name in [(pkg.TagFamily.T25H9, "t25h9"), (pkg.TagFamily.T16H5, "t16h5"),
                      (pkg.TagFamily.T25H7, "t25h7"), (pkg.TagFamily.T36H11B1, "t36h11b1")]:
        img = synth.render_board_numpy(640, 480, cols=5, rows=4, seed=9, tag_px=52.0,
                                       family=name)
        det = pkg.TagDetector(fam)
        try:
            got = det.detect(img)
        finally:
            det.close()
        want = oracle.detect(img, family=name)
        assert_tags_match(got, want)
        assert len(want) >= 15, (name, len(want))


# ---- board search variants: the four-lane group scoring path vs the general warp-wide path ----
BOARD_VARIANTS = [dict(board_fast=1, board_warps=1), dict(board_fast=1, board_warps=4),
                  dict(board_fast=1, board_warps=8), dict(board_fast=1, board_warps=16),
                  dict(board_fast=0, board_warps=1), dict(board_fast=0, board_warps=4),
                  dict(board_fast=0, board_warps=16)]


@pytest.mark.parametrize("variant", BOARD_VARIANTS, ids=lambda v: "fast%d_w%d" % (v["board_fast"], v["board_warps"]))
def test_board_search_variants_agree_with_oracle(pkg, oracle, images, variant):
    """Every mapping of the board search onto warps gives the oracle's quads, ids and corners."""
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        for k, v in variant.items():
            det.set_option(k, v)
        for name in ("EuRoC", "TUM_VI", "r45", "two_boards"):
            img = images[name]
            g = det.stages(img)
            o = oracle.front_end(img, want_labels=False)
            oq = oracle.try_find_best_board(o["refined"])
            assert oq is not None and np.array_equal(g["quads"], oq), name
            assert_tags_match(g["tags"], oracle.detect(img))
        frames = synth.fixture_like_frames(6, 640, 480, seed=40, tag_px=41.0)
        frames[2] = 0
        for gt, wt in zip(det.detect_batch(frames), oracle.detect_batch(frames)):
            assert_tags_match(gt, wt)
    finally:
        det.close()


def test_group_board_overflow_falls_back_to_general_build(pkg, oracle):
    """A board wider than the 16 x 16 group window is re-scored by the general path: a 14 x 2
    strip (28 tags, ~216 saddles, so the frame does take the throughput path)."""
    img = synth.render_board_numpy(1280, 1024, cols=14, rows=2, seed=4, tag_px=60.0)
    fe = oracle.front_end(img, want_labels=False)
    want = oracle.detect(img)
    assert len(want) == 28 and len(fe["refined"]) <= 512
    for fast, warps in ((1, 1), (1, 4), (0, 4)):
        det = pkg.TagDetector(pkg.TagFamily.T36H11)
        try:
            det.set_option("board_fast", fast)
            det.set_option("board_warps", warps)
            assert_tags_match(det.detect(img), want)
        finally:
            det.close()


def test_saddle_tiers_of_the_board_kernel(pkg, oracle):
    """809 refined saddles (12 x 9 board): with the 512-saddle tier the frame takes the general
    board path, with the 1024-saddle tier the four-lane group path; 108 tags either way, and the
    board is far wider than a group's 16 x 16 window (general re-scoring inside the group path)."""
    img = synth.render_board_numpy(1280, 1024, cols=12, rows=9, seed=4, tag_px=58.0)
    fe = oracle.front_end(img, want_labels=False)
    assert 512 < len(fe["refined"]) <= 1024
    want = oracle.detect(img)
    assert len(want) >= 100
    for tier in (0, 1):
        for warps in (2, 8, 16):
            det = pkg.TagDetector(pkg.TagFamily.T36H11)
            try:
                det.set_option("board_saddle_tier", tier)
                det.set_option("board_warps", warps)
                assert_tags_match(det.detect(img, cap=256), want)
            finally:
                det.close()


def test_4k_batch_takes_every_tier_of_the_board_kernel(pkg, oracle):
    """A batch of 4K frames in batch mode (three launches of the board kernel): a sparse 6 x 6 board
    (<= 320 saddles: small tier, throughput path), the 24 x 13 board of configs[3] and a 29 x 20 board
    of 580 tags (4096 tier: saddle list and 32 px bucket grid on chip, general path).  With the
    on-chip tier forced down to 1024 the two large frames keep their saddles and grid items in
    global memory.  Each equals the oracle every time."""
    frames = [synth.render_board_numpy(3840, 2160, cols=6, rows=6, seed=5, tag_px=110.0, ss=2),
              synth.render_board_numpy(3840, 2160, cols=24, rows=13, seed=3, tag_px=100.0, ss=2),
              synth.render_board_numpy(3840, 2160, cols=29, rows=20, seed=9, tag_px=84.0, ss=2)]
    n_ref = [len(oracle.front_end(f, want_labels=False)["refined"]) for f in frames]
    assert n_ref[0] <= 320 and 1024 < n_ref[1] <= 4096 and 3500 < n_ref[2] <= 4096, n_ref
    want = [oracle.detect(f, cap=1024) for f in frames]
    assert len(want[0]) == 36 and len(want[1]) > 290 and len(want[2]) > 500, [len(w) for w in want]
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        det.set_option("board_batch_frames", 1)  # batch mode (2 warps per frame, launches split by tier)
        tags, status = det.detect_batch(np.stack(frames), cap_per_frame=640, return_status=True)
        assert not status.any(), status
        for got, w in zip(tags, want):
            assert_tags_match(got, w)
        det.set_option("board_saddle_tier", 1)  # 1024 on chip: the large frames live in global memory
        tags, status = det.detect_batch(np.stack(frames), cap_per_frame=640, return_status=True)
        assert not status.any(), status
        for got, w in zip(tags, want):
            assert_tags_match(got, w)
        det.set_option("board_saddle_tier", -1)
        det.set_option("board_split", 0)  # one launch, one tier for all three
        tags, status = det.detect_batch(np.stack(frames), cap_per_frame=640, return_status=True)
        assert not status.any(), status
        for got, w in zip(tags, want):
            assert_tags_match(got, w)
    finally:
        det.close()


def test_4k_rgb_dense_board_through_detect_kornia(pkg, oracle):
    """BASELINE.json configs[3]: 3840 x 2160 RGB8, 24 x 13 tags, through detect / detect_kornia with
    DEFAULT options.  The frame has more refined saddles than a 1280 x 1024 frame's capacity (the
    automatic capacity follows the image area), far more than the on-chip tiers (general board
    path), more runs than the run-based labeller takes (pixel-list labeller), and the RGB rows go
    through the streaming K1.  With the capacity forced down to 2048 the host entry points re-run
    the frame with grown capacities (same result); the device entry point flags it."""
    import torch
    img = synth.render_board_numpy(3840, 2160, cols=24, rows=13, seed=3, tag_px=100.0, ss=2, rgb=True)
    fe = oracle.front_end(img, want_labels=False)
    want = oracle.detect(img)
    assert len(fe["refined"]) > 2048 and len(want) > 290
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        assert_tags_match(det.detect_kornia(img), want)
        assert_tags_match(det.detect(img), want)
        tags, status = det.detect_batch(img[None], cap_per_frame=512, return_status=True)
        assert status[0] == 0
        assert_tags_match(tags[0], want)
        det.set_option("max_saddles", 2048)  # too small for this frame
        tags, status = det.detect_batch(img[None], cap_per_frame=512, return_status=True)
        assert status[0] == 0  # re-run with a grown capacity inside the call
        assert_tags_match(tags[0], want)
        assert_tags_match(det.detect(img), want)
        d_img = torch.from_numpy(img).cuda()
        d_tags = torch.zeros((1, 512 * 9), dtype=torch.int32, device="cuda")
        d_cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        d_st = torch.zeros(1, dtype=torch.int32, device="cuda")
        det.detect_batch_device(d_img.data_ptr(), 1, 3840, 2160, pkg.FMT_RGB8, d_tags.data_ptr(), 512,
                                d_cnt.data_ptr(), d_st.data_ptr())
        torch.cuda.synchronize()
        assert int(d_st[0]) & 2  # AG_FRAME_SADDLE_OVERFLOW: device-resident results are flagged, not silent
    finally:
        det.close()


def test_detect_grows_capacities_instead_of_truncating(pkg, oracle):
    """The reference has no per-frame limits (detector.rs:505-540).  Frames that overflow the
    cluster / saddle capacities (iid noise: 65 k clusters, 11 k refined saddles; a fine checkerboard:
    6.9 k refined saddles) come back from detect / detect_batch / refined_saddle_points exactly as
    the oracle has them, with default options; a frame beyond the hard limits is an error, never a
    silently truncated map."""
    rng = np.random.default_rng(5)
    H, W = 1024, 1280
    noise = rng.integers(0, 256, (H, W), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    checker = np.where(((yy // 12) + (xx // 12)) % 2 == 0, 40, 200).astype(np.uint8)
    board = synth.render_board_numpy(W, H, seed=5)
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        for img, n_min in ((noise, 8000), (checker, 4000)):
            fe = oracle.front_end(img, want_labels=False)
            assert len(fe["refined"]) > n_min
            assert_saddles_match(det.refined_saddle_points(img, cap=32768), fe["refined"])
        assert_tags_match(det.detect(noise), oracle.detect(noise))
        frames = np.stack([board, noise, board, noise[::-1].copy()])
        want = [oracle.detect(f) for f in frames]
        assert len(want[0]) == 36
        tags, status = det.detect_batch(frames, return_status=True)
        assert not status.any()
        for t, w in zip(tags, want):
            assert_tags_match(t, w)
        # streaming calls: overflowed frames are re-run when their chunk is collected
        det.set_option("host_async", 1)
        out = np.zeros((4, 128), pkg.TAG_DTYPE)
        cnt = np.zeros(4, np.int32)
        st = np.ones(4, np.uint32)
        det.detect_batch_into(frames, out, cnt, st)
        det.detect_batch_wait(0)
        det.set_option("host_async", 0)
        assert not st.any() and cnt.tolist() == [len(w) for w in want]
        # beyond the hard limit of 16384 saddles per frame: an error, not a truncated result
        dense = np.where(((yy // 5) + (xx // 5)) % 2 == 0, 40, 200).astype(np.uint8)
        if len(oracle.front_end(dense, want_labels=False)["refined"]) > 16384:
            with pytest.raises(RuntimeError, match="limits"):
                det.detect(dense)
            with pytest.raises(RuntimeError, match="limits"):
                det.refined_saddle_points(dense, cap=65536)
    finally:
        det.close()


def test_detect_is_synchronous_on_a_streaming_handle(pkg, oracle):
    """ag_detect writes into buffers that usually live on the caller's stack: it must return complete
    results even when the handle's streaming option (host_async) is on, and a device-batch call on
    the same handle must not overtake host chunks that are still in flight."""
    import torch
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        n, w, h, cap = 24, 640, 480, 64
        det.set_option("host_chunk_frames", 4)
        d_frames = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        det.render_boards_device(d_frames.data_ptr(), n, w, h, 6, 6, 4242)
        torch.cuda.synchronize()
        frames = d_frames.cpu().numpy()
        ref = (np.zeros((n, cap), pkg.TAG_DTYPE), np.zeros(n, np.int32), np.zeros(n, np.uint32))
        det.detect_batch_into(frames, *ref)
        img = synth.render_board_numpy(w, h, seed=9, tag_px=44.0)
        want = oracle.detect(img)
        assert len(want) == 36
        det.set_option("host_async", 1)
        for rep in range(3):
            o = (np.zeros((n, cap), pkg.TAG_DTYPE), np.zeros(n, np.int32), np.zeros(n, np.uint32))
            det.detect_batch_into(frames, *o)          # chunks in flight
            assert_tags_match(det.detect(img), want)   # synchronous nevertheless, and correct
            assert np.array_equal(o[0], ref[0]) and np.array_equal(o[1], ref[1])  # ... and it delivered them
            det.detect_batch_wait(0)
        # device call while host chunks are in flight
        o = (np.zeros((n, cap), pkg.TAG_DTYPE), np.zeros(n, np.int32), np.zeros(n, np.uint32))
        det.detect_batch_into(frames, *o)
        tags = torch.zeros((n, cap * 9), dtype=torch.int32, device="cuda")
        cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
        det.detect_batch_device(d_frames.data_ptr(), n, w, h, pkg.FMT_L8, tags.data_ptr(), cap, cnt.data_ptr())
        torch.cuda.synchronize()
        det.detect_batch_wait(0)
        assert np.array_equal(o[0], ref[0]) and np.array_equal(o[1], ref[1])
        assert np.array_equal(cnt.cpu().numpy(), ref[1])
        assert np.array_equal(tags.cpu().numpy().view(pkg.TAG_DTYPE).reshape(n, cap), ref[0])
        det.set_option("host_async", 0)
    finally:
        det.close()


def test_label_kernel_variants_agree(pkg, oracle, images):
    """K3 run-based (shared-memory union-find over runs) and pixel-list versions: same cluster
    centres, same order, bit for bit, on images, noise (fallback: too many runs) and odd shapes."""
    rng = np.random.default_rng(3)
    cases = [images["EuRoC"], images["two_boards"], images["TUM_VI"],
             rng.integers(0, 256, (200, 333), dtype=np.uint8),
             rng.integers(0, 256, (64, 40), dtype=np.uint8),
             synth.render_board_numpy(640, 480, seed=8, tag_px=44.0)]
    # long horizontal structures: runs spanning several mask words
    stripes = np.zeros((96, 400), np.uint8)
    stripes[::7] = 255
    stripes[:, ::53] = 128
    cases.append(stripes)
    dets = []
    try:
        for variant in (0, 1):
            d = pkg.TagDetector(pkg.TagFamily.T36H11)
            d.set_option("label_variant", variant)
            d.set_option("max_clusters", 1 << 17)
            dets.append(d)
        for img in cases:
            a, b = dets[0].stages(img), dets[1].stages(img)
            o = oracle.front_end(img)
            assert np.array_equal(a["centers"].view(np.uint32), b["centers"].view(np.uint32))
            assert np.array_equal(a["labels"], o["labels"]) and np.array_equal(b["labels"], o["labels"])
            if len(o["centers"]) == len(a["centers"]):
                differ = (a["centers"].view(np.uint32) != o["centers"].view(np.uint32)).any(axis=1)
                sizes = np.bincount(o["labels"][o["labels"] >= 0], minlength=len(differ))
                assert (sizes[differ] * max(img.shape[:2]) >= 2 ** 24).all()  # documented deviation only
    finally:
        for d in dets:
            d.close()


@pytest.mark.parametrize("kw", [dict(dtype=np.uint16), dict(rgb=True)])
def test_detect_batch_other_formats(pkg, oracle, kw):
    """16-bit gray and RGB frames through the batched host path (ragged chunks)."""
    frames = np.stack([synth.render_board_numpy(640, 480, seed=60 + i, tag_px=42.0, **kw) for i in range(5)])
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        det.set_option("host_chunk_frames", 2)
        got = det.detect_batch(frames)
    finally:
        det.close()
    want = oracle.detect_batch(frames)
    assert len(got) == 5
    for g, w in zip(got, want):
        assert len(w) == 36
        assert_tags_match(g, w)


def test_streaming_host_calls_equal_synchronous_calls(pkg):
    """ag_detect_batch with host_async + ag_detect_batch_wait: calls overlap (a later call recycles the
    staging of an earlier one and hands its results out), different batches and capacities in
    flight; results are byte-identical to synchronous calls; AG_ERR_CAPACITY arrives at the wait."""
    import torch
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        n, w, h, cap = 44, 640, 480, 64
        det.set_option("host_chunk_frames", 4)  # 11 chunks per call: more than the 8 board slots
        d_frames = torch.empty((2 * n, h, w), dtype=torch.uint8, device="cuda")
        det.render_boards_device(d_frames.data_ptr(), 2 * n, w, h, 6, 6, 77)
        torch.cuda.synchronize()
        frames = d_frames.cpu().numpy()
        batches = [frames[:n], frames[n:], frames[5:n + 5]]

        def new_out(c):
            return (np.zeros((n, c), pkg.TAG_DTYPE), np.zeros(n, np.int32), np.zeros(n, np.uint32))

        ref = []
        for b in batches:
            o = new_out(cap)
            det.detect_batch_into(b, *o)
            ref.append(o)
        assert ref[0][1].sum() > 0 and not np.array_equal(ref[0][0], ref[1][0])
        det.set_option("host_async", 1)
        outs = [new_out(cap) for _ in batches]
        for k, b in enumerate(batches):
            det.detect_batch_into(b, *outs[k])
            if k == 1:
                det.detect_batch_wait(1)  # call 0 is complete, call 1 may still be in flight
                assert np.array_equal(outs[0][1], ref[0][1]) and np.array_equal(outs[0][0], ref[0][0])
        det.detect_batch_wait(0)
        for o, r in zip(outs, ref):
            assert np.array_equal(o[1], r[1]) and np.array_equal(o[0], r[0]) and np.array_equal(o[2], r[2])
        # capacity overflow of a streaming call is reported by the wait; counts stay exact
        small = new_out(2)
        det.detect_batch_into(batches[0], *small)
        with pytest.raises(RuntimeError, match="cap_per_frame"):
            det.detect_batch_wait(0)
        assert np.array_equal(small[1], ref[0][1])
        assert np.array_equal(small[0], ref[0][0][:, :2])
        det.detect_batch_wait(0)  # nothing pending, no error left over
        # back to synchronous calls, and the device path after streaming host calls
        det.detect_batch_into(batches[1], *outs[0])
        det.set_option("host_async", 0)
        o = new_out(cap)
        det.detect_batch_into(batches[2], *o)
        assert np.array_equal(outs[0][0], ref[1][0]) and np.array_equal(o[0], ref[2][0])
    finally:
        det.close()


def test_streaming_device_calls_equal_synchronous_calls(pkg, oracle):
    """ag_detect_batch_device with device_async + ag_detect_batch_device_wait: several calls in flight
    over all board slots give byte-identical results to one synchronising call per batch."""
    import torch
    det = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        n, w, h, cap = 80, 640, 480, 64
        det.set_option("chunk_frames", 8)  # 10 chunks per call: more than the 8 board slots
        frames = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        det.render_boards_device(frames.data_ptr(), n, w, h, 6, 6, 2024)
        torch.cuda.synchronize()

        def run(k_calls, async_mode):
            det.set_option("device_async", 1 if async_mode else 0)
            outs = []
            s = torch.cuda.current_stream().cuda_stream
            for _ in range(k_calls):
                tags = torch.zeros((n, cap * 9), dtype=torch.int32, device="cuda")
                cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
                st = torch.zeros(n, dtype=torch.int32, device="cuda")
                det.detect_batch_device(frames.data_ptr(), n, w, h, pkg.FMT_L8, tags.data_ptr(), cap,
                                        cnt.data_ptr(), st.data_ptr(), stream=s)
                outs.append((tags, cnt, st))
            if async_mode:
                det.detect_batch_device_wait(stream=s)
            torch.cuda.synchronize()
            det.set_option("device_async", 0)
            return outs

        ref = run(1, False)[0]
        for tags, cnt, st in run(3, True):
            assert torch.equal(cnt, ref[1]) and torch.equal(tags, ref[0]) and int(st.abs().sum()) == 0
        # and the batch agrees with the oracle on a sample
        host = frames[:4].cpu().numpy()
        want = oracle.detect_batch(host)
        rec = ref[0][:4].cpu().numpy().view(pkg.TAG_DTYPE).reshape(4, cap)
        cn = ref[1][:4].cpu().numpy()
        for i in range(4):
            got = {int(t["id"]): t["xy"].reshape(4, 2) for t in rec[i, :cn[i]]}
            assert_tags_match(got, want[i])
    finally:
        det.close()


def test_label_kernel_row_limit_fallback(detector, oracle):
    """More rows than the run-based labelling kernel keeps row starts for (4096): the frame is handed
    to the pixel-list kernel; stages still match the oracle."""
    rng = np.random.default_rng(21)
    img = np.full((4200, 48), 120, np.uint8)
    img[::9, ::5] = 20
    img[3::17] = 230
    img = (img.astype(np.int32) + rng.integers(-3, 4, img.shape)).clip(0, 255).astype(np.uint8)
    check_stages(detector, oracle, img, check_board=False)


def test_multi_gpu_detect_batch_equals_single_gpu(pkg):
    """ag_multi_detect_batch: the batch is sharded image-wise over the GPUs of the box, one host thread
    per device; records come back in frame order, byte-identical to one GPU doing the whole batch.
    With one visible GPU the same device is used twice (two handles, two shards)."""
    import torch
    n_gpu = torch.cuda.device_count()
    devices = list(range(n_gpu)) if n_gpu >= 2 else [0, 0]
    single = pkg.TagDetector(pkg.TagFamily.T36H11, None, device=0)
    multi = pkg.MultiTagDetector(pkg.TagFamily.T36H11, None, devices=devices)
    try:
        assert multi.n_devices == len(devices)
        n, w, h, cap = 37, 640, 480, 64  # 37: ragged shards
        d_frames = torch.empty((n, h, w), dtype=torch.uint8, device="cuda:0")
        single.render_boards_device(d_frames.data_ptr(), n, w, h, 6, 6, 9001)
        torch.cuda.synchronize()
        frames = d_frames.cpu().numpy()
        ref = (np.zeros((n, cap), pkg.TAG_DTYPE), np.zeros(n, np.int32), np.zeros(n, np.uint32))
        single.detect_batch_into(frames, *ref)
        got = (np.zeros((n, cap), pkg.TAG_DTYPE), np.zeros(n, np.int32), np.ones(n, np.uint32))
        multi.set_option("host_chunk_frames", 5)
        multi.detect_batch_into(frames, *got)
        assert ref[1].sum() > 30 * n
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2])
        assert len(multi.detect_batch(frames[:3])) == 3 and multi.detect_batch(frames[:0]) == []
        small = (np.zeros((n, 2), pkg.TAG_DTYPE), np.zeros(n, np.int32), np.zeros(n, np.uint32))
        with pytest.raises(RuntimeError, match="cap_per_frame"):
            multi.detect_batch_into(frames, *small)
        assert np.array_equal(small[1], ref[1])
    finally:
        multi.close()
        single.close()


def test_cpp_mirror_end_to_end(pkg, oracle, tmp_path):
    """The C++ mirror of the reference API (cpp/aprilgrid_b200.hpp), compiled and run: detect,
    detect_batch and MultiTagDetector give the oracle's tags."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cpp_mirror_demo")
    libdir = os.path.join(root, "aprilgrid-rs_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "aprilgrid-rs_b200", "cpp"), "-o", exe,
                           os.path.join(root, "tests", "cpp_mirror_demo.cpp"), "-L" + libdir, "-laprilgrid_b200",
                           "-Wl,-rpath," + libdir])
    img = synth.render_board_numpy(640, 480, seed=12, tag_px=44.0)
    raw = tmp_path / "frame.raw"
    raw.write_bytes(img.tobytes())
    out = subprocess.check_output([exe, str(raw), "640", "480"], text=True).strip().splitlines()
    want = oracle.detect(img)
    assert out[-1].startswith("batch same")
    got = {}
    for ln in out[:-1]:
        f = ln.split()
        got[int(f[0])] = np.array([float(v) for v in f[1:]], np.float32).reshape(4, 2)
    assert_tags_match(got, want)


def test_device_renderer_geometry_is_the_chart_generators(detector, pkg, oracle):
    """The on-device renderer against the chart definition (scripts/generate_aprilgrid.py:1114-1167)
    and against tests/synth.py: under a GIVEN pose, the noiseless device frame equals the numpy
    rendering pixel for pixel (up to f32 edge cases), and tag id i + j * cols sits at lattice column
    i, row j counted from the BOTTOM of the board, its four corners where the pose puts the corners
    of that tag."""
    import torch
    w, h, cols, rows = 960, 720, 6, 5
    scale, tx, ty = 80.0, 90.0, 60.0  # frontal view: page (X, Y) in tag sides -> image px
    H = np.array([[scale, 0.0, tx], [0.0, scale, ty], [0.0, 0.0, 1.0]])
    d = torch.empty((h, w), dtype=torch.uint8, device="cuda")
    detector._render_pose(d.data_ptr(), w, h, cols, rows, np.linalg.inv(H), noise=False)
    got = d.cpu().numpy()
    want = synth.render_board_numpy(w, h, cols, rows, H=H, noise=0.0)
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert (diff > 0).mean() < 1e-3 and diff.max() <= 11  # a 1/16 coverage step at the odd f32 edge sample
    tags = detector.detect(got)
    assert sorted(tags) == list(range(cols * rows))
    assert_tags_match(tags, oracle.detect(got))
    hb = rows * 1.3 + 0.3
    for tid, xy in tags.items():
        i, j = tid % cols, tid // cols              # ids run row-major from the bottom row
        x0, x1 = 0.3 + 1.3 * i, 1.3 * (i + 1)        # page extent of the tag, Y measured from the top
        y1, y0 = hb - (0.3 + 1.3 * j), hb - 1.3 * (j + 1)
        # integer image coordinates are pixel centres in the renderer and in the detector alike
        exp = {(scale * X + tx, scale * Y + ty) for X in (x0, x1) for Y in (y0, y1)}
        for cx, cy in xy:
            assert min(abs(cx - ex) + abs(cy - ey) for ex, ey in exp) < 0.6, (tid, xy, sorted(exp))
        assert len({(int(round(cx)), int(round(cy))) for cx, cy in xy}) == 4


def test_board_search_on_gate_adversarial_saddles(detector, oracle):
    """GPU board search == oracle on saddle lists constructed to sit on every discontinuous gate (a
    1-ulp disagreement in an angle would flip a candidate there).  Quads of the first best board,
    same set, same order."""
    img = synth.render_board_numpy(640, 480, seed=3, tag_px=44.0)
    base = oracle.front_end(img, want_labels=False)["refined"]
    assert len(base) > 150
    rng = np.random.default_rng(2024)
    n_boards = 0
    for name, s in synth.adversarial_saddle_sets(base, rng):
        want = oracle.try_find_best_board(s)
        got, _ = detector._boards_from_saddles(s, img)
        if want is None:
            assert len(got) == 0, name
        else:
            assert np.array_equal(got, want), name
            n_boards += 1
    assert n_boards >= 30


def test_detect_planes_matches_oracle(detector, oracle):
    """ag_detect_planes: the frame as its two derived gray planes (to_luma32f / to_luma8), for
    DynamicImage variants whose conversion the caller does with `image` itself.  Planes that no
    single L8 / L16 / RGB8 image would produce (float luma from a 16-bit source with its own
    rounding, an independently rounded 8-bit plane): tags identical to the oracle on the same planes,
    and identical to plain detect when the planes ARE an L8 image's own."""
    img16 = synth.render_board_numpy(800, 600, seed=21, tag_px=52.0, dtype=np.uint16)
    rng = np.random.default_rng(4)
    jitter = rng.integers(-90, 91, img16.shape)
    v = np.clip(img16.astype(np.int64) + jitter, 0, 65535)
    luma32f = (v.astype(np.float64) * (0.2126 + 0.7152 + 0.0722) / 65535.0).astype(np.float32)
    luma8 = ((v + 77) // 257).clip(0, 255).astype(np.uint8)
    want = oracle.detect_planes(luma32f, luma8)
    assert len(want) == 36
    assert_tags_match(detector.detect_planes(luma32f, luma8), want)
    img8 = synth.render_board_numpy(640, 480, seed=3, tag_px=44.0)
    got = detector.detect_planes(oracle.to_luma_f32(img8), oracle.to_luma_u8(img8))
    ref = detector.detect(img8)
    assert sorted(got) == sorted(ref) and all(np.array_equal(got[k], ref[k]) for k in ref)
    # a width that is not a multiple of four takes the tile kernel's f32 instantiation
    img_odd = synth.render_board_numpy(642, 481, seed=5, tag_px=44.0)
    f_odd, u_odd = oracle.to_luma_f32(img_odd), oracle.to_luma_u8(img_odd)
    want_odd = oracle.detect_planes(f_odd, u_odd)
    assert len(want_odd) == 36
    assert_tags_match(detector.detect_planes(f_odd, u_odd), want_odd)
