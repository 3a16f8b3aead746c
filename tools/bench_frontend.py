"""Image front-end throughput: PIL on the host cores (what the reference does, ThreadPool(8)) vs the device resampler,
for scan-sized pages. Prints pages/s for each and the kernel-only time of the device path.
  python tools/bench_frontend.py [--pages 64] [--size 1700x2200]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from visrag_b200.config import VisRAGConfig  # noqa: E402
from visrag_b200.frontend import DeviceFrontEnd  # noqa: E402
from visrag_b200.host import plan_slices, render_slices  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=64)
    ap.add_argument("--size", default="1700x2200")
    a = ap.parse_args()
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    W, H = (int(x) for x in a.size.split("x"))
    cfg = VisRAGConfig.full()
    plan = plan_slices(W, H, cfg)
    rs = np.random.RandomState(0)
    arr = rs.randint(0, 256, (a.pages, H, W, 3), dtype=np.uint8)
    imgs = [Image.fromarray(x) for x in arr]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=8) as ex:
        host = list(ex.map(lambda im: render_slices(im, plan), imgs))
    t_host = time.perf_counter() - t0
    print(f"page {W}x{H}: plan thumb {plan.source_size} grid {plan.grid} refine {plan.refine_size} cell {plan.cell_size}")
    print(f"host PIL, 8 threads: {a.pages / t_host:8.1f} pages/s ({t_host / a.pages * 1e3:.2f} ms/page)")

    fe = DeviceFrontEnd(torch.device("cuda:0"))
    pinned = torch.from_numpy(arr).pin_memory()
    tw, th = plan.source_size
    thumbs = torch.empty((a.pages, th, tw, 3), dtype=torch.uint8, device="cuda:0")
    first_t = torch.arange(a.pages, dtype=torch.int32, device="cuda:0")
    cells = None
    if plan.grid is not None:
        cw, ch = plan.cell_size
        nc = plan.grid[0] * plan.grid[1]
        cells = torch.empty((a.pages * nc, ch, cw, 3), dtype=torch.uint8, device="cuda:0")
        first_c = torch.arange(0, a.pages * nc, nc, dtype=torch.int32, device="cuda:0")

    def device_pass(dev_pages):
        fe.resize_into(dev_pages, tw, th, thumbs, first_t, tw, th)
        if cells is not None:
            fe.resize_into(dev_pages, plan.refine_size[0], plan.refine_size[1], cells, first_c, cw, ch)

    dev_pages = pinned.cuda(non_blocking=True)
    for _ in range(2):
        device_pass(dev_pages)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 5
    e0.record()
    for _ in range(reps):
        dev_pages = pinned.cuda(non_blocking=True)
    e1.record()
    for _ in range(reps):
        device_pass(dev_pages)
    e2.record()
    torch.cuda.synchronize()
    t_copy, t_kern = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
    src_bytes = arr.nbytes
    print(f"device: H2D {t_copy:.2f} ms ({src_bytes / t_copy / 1e6:.1f} GB/s)  kernels {t_kern:.2f} ms "
          f"({src_bytes / t_kern / 1e6:.1f} GB/s of source pixels)  -> {a.pages / (t_copy + t_kern) * 1e3:8.1f} pages/s")
    # exactness on this shape
    got_t = thumbs.cpu().numpy()
    ok = all(np.array_equal(got_t[i], host[i][0]) for i in range(a.pages))
    if cells is not None:
        got_c = cells.cpu().numpy()
        ok &= all(np.array_equal(got_c[i * nc + c], host[i][1 + c]) for i in range(a.pages) for c in range(nc))
    print("bit-identical to PIL:", ok)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
