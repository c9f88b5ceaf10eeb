import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import __graft_entry__ as entry
pkg = entry.load_package()
det = pkg.TagDetector(pkg.TagFamily.T36H11)
n, W, H = 256, 1280, 1024
for name, fmt, shape, dt in (("L8", pkg.FMT_L8, (n, H, W), torch.uint8), ("L16", pkg.FMT_L16, (n, H, W), torch.int16), ("RGB8", pkg.FMT_RGB8, (n, H, W, 3), torch.uint8)):
    fr = torch.randint(0, 127, shape, dtype=dt, device="cuda")
    for variant in (0, 1):
        det.set_option("dense_variant", variant)
        for _ in range(2): det.dense_batch_device(fr.data_ptr(), n, W, H, fmt)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): det.dense_batch_device(fr.data_ptr(), n, W, H, fmt)
        torch.cuda.synchronize(); dt_ = (time.perf_counter() - t0) / 5
        print("%s variant %d: %.3f ms per %d frames = %.0f frames/s (K1+K2)" % (name, variant, dt_ * 1e3, n, n / dt_))
det.close()
