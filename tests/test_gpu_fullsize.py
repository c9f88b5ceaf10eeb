"""BASELINE.json's full sizes (1280x1024 batches) through the device-resident C-ABI entry points.
Per-frame oracle comparison on a sample, size-independent properties on the whole batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 1280, 1024


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


def render(det, torch, n, seed, w=W, h=H):
    frames = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    det.render_boards_device(frames.data_ptr(), n, w, h, 6, 6, seed)
    torch.cuda.synchronize()
    return frames


def run_device(det, pkg, torch, frames, cap=64):
    n, h, w = frames.shape
    out = torch.zeros((n, cap * 9), dtype=torch.int32, device="cuda")  # ag_tag = 36 bytes
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    status = torch.zeros(n, dtype=torch.int32, device="cuda")
    det.detect_batch_device(frames.data_ptr(), n, w, h, pkg.FMT_L8, out.data_ptr(), cap, cnt.data_ptr(),
                            status.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    rec = out.cpu().numpy().view(pkg.TAG_DTYPE).reshape(n, cap)
    return rec, cnt.cpu().numpy(), status.cpu().numpy()


def test_batch_1280x1024_properties_and_sampled_oracle(detector, pkg, oracle, torch_mod):
    torch = torch_mod
    n = 96
    frames = render(detector, torch, n, seed=1234)
    rec, cnt, status = run_device(detector, pkg, torch, frames)
    assert (status == 0).all()
    # nearly every rendered 6x6 board decodes to ids 0..35 exactly once; the odd frame whose pose
    # defeats the detector must defeat the oracle in exactly the same way (checked below)
    full = cnt == 36
    assert full.mean() >= 0.9, cnt
    for i in np.nonzero(full)[0]:
        assert list(rec[i, :36]["id"]) == list(range(36))
    assert (cnt <= 36).all()
    # idempotence: same input, same bytes out
    rec2, cnt2, _ = run_device(detector, pkg, torch, frames)
    assert np.array_equal(cnt, cnt2) and np.array_equal(rec.tobytes(), rec2.tobytes())
    # batch-composition independence: a frame's result does not depend on its neighbours
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    rec3, cnt3, _ = run_device(detector, pkg, torch, frames[perm].contiguous())
    p = perm.cpu().numpy()
    assert np.array_equal(rec3.tobytes(), rec[p].tobytes())
    # sampled frames against the oracle
    sample = sorted(set(range(0, n, 16)) | set(int(i) for i in np.nonzero(~full)[0]))
    host = frames[sample].cpu().numpy()
    want = oracle.detect_batch(host, threads=8)
    for i, wtags in zip(sample, want):
        got = {int(t["id"]): t["xy"].reshape(4, 2) for t in rec[i, :cnt[i]]}
        assert sorted(got) == sorted(wtags)
        for k in wtags:
            assert np.abs(got[k] - wtags[k]).max() <= 1e-3  # px, north_star tolerance


def test_dense_batch_device_stage_parity_at_full_size(detector, pkg, oracle, torch_mod):
    """The blur/threshold configuration: one 1280x1024 noise frame and one board frame."""
    torch = torch_mod
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (H, W), dtype=np.uint8)
    from test_gpu_parity import check_stages
    check_stages(detector, oracle, noise, check_board=False)  # > 16384 clusters: truncated + flagged
    big = pkg.TagDetector(pkg.TagFamily.T36H11)
    try:
        big.set_option("max_clusters", 1 << 18)
        big.set_option("max_saddles", 16384)
        g, o = check_stages(big, oracle, noise, check_board=False)
        assert len(g["centers"]) == len(o["centers"]) > 16384
    finally:
        big.close()
    board = render(detector, torch, 1, seed=99)[0].cpu().numpy()
    check_stages(detector, oracle, board)


def test_large_rgb_frame(detector, pkg, oracle, torch_mod):
    """4K RGB through detect (the detect_kornia configuration), dense 12x7 board."""
    torch = torch_mod
    w, h = 3840, 2160
    gray = torch.empty((1, h, w), dtype=torch.uint8, device="cuda")
    detector.render_boards_device(gray.data_ptr(), 1, w, h, 12, 7, 4321)
    torch.cuda.synchronize()
    g = gray[0].cpu().numpy()
    rgb = np.repeat(g[:, :, None], 3, axis=2).copy()
    got = detector.detect(rgb)
    want = oracle.detect(rgb)
    assert sorted(got) == sorted(want) and len(want) >= 60
    for k in want:
        assert np.abs(got[k] - want[k]).max() <= 1e-3
