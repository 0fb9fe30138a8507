"""Golden vectors for the any-res preprocessing: the REAL reference function (omchat/mm_utils.py process_anyres_image) with
the CLIPImageProcessor the reference builds (internVIT_encoder.py:26-29), run in this container on seeded synthetic
images. Run once: python tests/golden/make_golden_preprocess.py  ->  tests/golden/golden_preprocess.npz
(the reference cannot travel to the GPU box; the vectors can). The input images are regenerated from seeds
(tests/golden/preprocess_images.py); the fixture stores per-crop checksums and strided samples of every crop."""
import os
import sys
import types

import numpy as np

sys.path.insert(0, "/root/reference")
import transformers  # noqa: E402,F401

for name in ("timm", "timm.models", "timm.models.layers", "timm.layers", "timm.models.regnet", "peft", "accelerate"):
    sys.modules.setdefault(name, types.ModuleType(name))

import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_mm_utils", "/root/reference/omchat/mm_utils.py")
constants = types.ModuleType("omchat.constants")
constants.IMAGE_TOKEN_INDEX = -200
sys.modules.setdefault("omchat", types.ModuleType("omchat"))
sys.modules["omchat.constants"] = constants
mm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mm)

from PIL import Image  # noqa: E402
from transformers import CLIPImageProcessor  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from preprocess_images import PINPOINTS, SIZES, synthetic_image  # noqa: E402


def main():
    proc = CLIPImageProcessor(crop_size=448, do_center_crop=True, do_normalize=True, do_resize=True,
                              image_mean=[0.485, 0.456, 0.406], image_std=[0.229, 0.224, 0.225], size=448)
    out = {}
    for n, (W, H) in enumerate(SIZES):
        img = synthetic_image(n, W, H)  # smooth + noisy content: exercises the filter taps and the rounding
        pixels, best = mm.process_anyres_image(Image.fromarray(img), proc, PINPOINTS, return_best_res=True)
        out[f"img{n}"] = img
        out[f"best{n}"] = np.asarray(best, dtype=np.int32)
        out[f"pix{n}"] = pixels.numpy().astype(np.float32)
        # keep the file small: store every crop's checksum and only two full crops per image
        print(n, (W, H), "best", best, "crops", tuple(pixels.shape))
    slim = {}
    for n in range(len(SIZES)):
        p = out[f"pix{n}"]
        slim[f"imgsum{n}"] = np.asarray([int(out[f"img{n}"].astype(np.int64).sum())], dtype=np.int64)
        slim[f"best{n}"] = out[f"best{n}"]
        slim[f"ncrops{n}"] = np.asarray([p.shape[0]], dtype=np.int32)
        slim[f"sum{n}"] = p.astype(np.float64).sum(axis=(1, 2, 3))
        slim[f"abs{n}"] = np.abs(p).astype(np.float64).sum(axis=(1, 2, 3))
        slim[f"samples{n}"] = p[:, :, 3::7, 5::7].copy()  # strided fp32 samples of EVERY crop: exact compare
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_preprocess.npz"), **slim)


if __name__ == "__main__":
    main()
