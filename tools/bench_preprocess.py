#!/usr/bin/env python
"""Any-res preprocessing throughput: GPU kernels (host uint8 image in pinned memory -> crops on the device, H2D inside the
timed region) vs the reference's CPU path restated with Pillow itself (Image.resize + paste + crop + numpy normalise) on the
host cores. Prints images/s for a few source sizes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402

from omchat_b200 import preprocess as PP  # noqa: E402
from preprocess_images import PINPOINTS, synthetic_image  # noqa: E402


def cpu_reference(pil, lut):
    W, H = pil.size
    best = PP.select_best_resolution((W, H), PINPOINTS)
    nw, nh, px, py = PP.resize_and_pad_geometry((W, H), best)
    canvas = Image.new("RGB", best, (0, 0, 0))
    canvas.paste(pil.resize((nw, nh)), (px, py))
    crops = [pil.resize((448, 448))] + [canvas.crop((j, i, j + 448, i + 448)) for i in range(0, best[1], 448)
                                        for j in range(0, best[0], 448)]
    out = np.empty((len(crops), 3, 448, 448), dtype=np.float32)
    for n, c in enumerate(crops):
        a = np.asarray(c)
        for ch in range(3):
            out[n, ch] = lut[ch][a[:, :, ch]]
    return out


def main():
    pre = PP.AnyResPreprocessor(PINPOINTS, dtype=torch.bfloat16)
    lut = PP.normalize_lut()
    for (W, H) in [(640, 480), (1920, 1080), (4032, 3024)]:
        img = synthetic_image(0, W, H)
        pinned = torch.from_numpy(img).pin_memory()
        pil = Image.fromarray(img)
        for _ in range(3):
            pre(pinned)
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            out = pre(pinned)
        torch.cuda.synchronize()
        gpu = reps / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 2.0:
            cpu_reference(pil, lut)
            n += 1
        cpu = n / (time.perf_counter() - t0)
        print(f"{W}x{H}: {tuple(out.shape)} crops  GPU {gpu:8.1f} img/s (H2D of {img.nbytes / 1e6:.1f} MB included)   "
              f"CPU Pillow path (1 thread) {cpu:6.1f} img/s   x{gpu / cpu:.1f}")


if __name__ == "__main__":
    main()
