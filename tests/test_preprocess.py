"""Any-resolution preprocessing.
CPU: the numpy oracle (oracle/preprocess_oracle.py) is pinned (1) bit for bit against Pillow's own Image.resize on random
images and (2) against vectors produced by the reference's process_anyres_image (tests/golden/make_golden_preprocess.py);
the product's vectorised coefficient tables equal the oracle's loop restatement.
GPU (-m gpu): omc_resample_u8 / omc_anyres_pack through AnyResPreprocessor equal the oracle exactly (uint8 pixels and fp32
crops), on the golden images and on random shapes."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as PO  # checker only
from omchat_b200 import preprocess as PP
from preprocess_images import PINPOINTS, SIZES, synthetic_image

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_preprocess.npz"))


def test_oracle_resize_is_pillow_bit_exact():
    from PIL import Image
    rs = np.random.RandomState(0)
    for (W, H, w, h) in [(640, 480, 448, 448), (123, 77, 448, 300), (1000, 700, 896, 627), (500, 1333, 336, 896),
                         (448, 448, 448, 448), (300, 448, 448, 448), (31, 17, 5, 3), (2, 2, 448, 448), (900, 1, 448, 7)]:
        a = rs.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
        want = np.asarray(Image.fromarray(a).resize((w, h)))
        assert np.array_equal(PO.resize(a, (w, h)), want), (W, H, w, h)


def test_oracle_matches_reference_golden():
    for n, (W, H) in enumerate(SIZES):
        img = synthetic_image(n, W, H)
        assert int(img.astype(np.int64).sum()) == int(GOLD[f"imgsum{n}"][0]), "synthetic image generator drifted"
        best = PO.select_best_resolution((W, H), PINPOINTS)
        assert tuple(best) == tuple(GOLD[f"best{n}"])
        out = PO.process_anyres(img, PINPOINTS)
        assert out.shape[0] == int(GOLD[f"ncrops{n}"][0])
        # pixels are exact; the float normalisation of the transformers-5.5 torchvision backend that generated the vectors
        # differs from the 4.41 numpy formulas restated here by at most one fp32 ulp of values <= 2.7
        assert np.abs(out[:, :, 3::7, 5::7] - GOLD[f"samples{n}"]).max() <= 2.4e-7
        assert np.abs(out.astype(np.float64).sum(axis=(1, 2, 3)) - GOLD[f"sum{n}"]).max() < 0.1


def test_product_tables_and_geometry_match_oracle():
    for (i, o) in [(640, 448), (123, 448), (1000, 896), (1333, 896), (300, 448), (2000, 1344), (97, 448), (5, 448), (448, 3)]:
        c1, b1 = PP.precompute_coeffs(i, o)
        c2, b2 = PO.precompute_coeffs(i, o)
        assert np.array_equal(c1, c2) and np.array_equal(b1, b2)
    assert np.array_equal(PP.normalize_lut(), PO.normalize_lut())
    rs = np.random.RandomState(1)
    for _ in range(200):
        size = (int(rs.randint(1, 3000)), int(rs.randint(1, 3000)))
        best = PP.select_best_resolution(size, PINPOINTS)
        assert best == PO.select_best_resolution(size, PINPOINTS)
        assert PP.resize_and_pad_geometry(size, best) == PO.resize_and_pad_geometry(size, best)
    assert PP.get_anyres_image_grid_shape((640, 480), str(PINPOINTS), 448) == (2, 2)


@pytest.mark.gpu
def test_gpu_preprocess_equals_oracle_exactly():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pre = PP.AnyResPreprocessor(PINPOINTS)
    for n, (W, H) in enumerate(SIZES):
        img = synthetic_image(n, W, H)
        got, best = pre(img, return_best_res=True)
        want = PO.process_anyres(img, PINPOINTS)
        assert tuple(best) == tuple(GOLD[f"best{n}"]) and tuple(got.shape) == want.shape
        assert torch.equal(got.cpu(), torch.from_numpy(want)), f"image {n}: fp32 crops differ"
        assert np.abs(got[:, :, 3::7, 5::7].cpu().numpy() - GOLD[f"samples{n}"]).max() <= 2.4e-7  # the reference's own output
    # bf16 output = rounding of the same values
    b = PP.AnyResPreprocessor(PINPOINTS, dtype=torch.bfloat16)(synthetic_image(0, *SIZES[0]))
    assert torch.equal(b.cpu(), torch.from_numpy(PO.process_anyres(synthetic_image(0, *SIZES[0]), PINPOINTS)).to(torch.bfloat16))


@pytest.mark.gpu
def test_gpu_resize_is_pillow_bit_exact():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from PIL import Image
    pre = PP.AnyResPreprocessor(PINPOINTS)
    rs = np.random.RandomState(3)
    for (W, H, w, h) in [(640, 480, 448, 448), (123, 77, 448, 300), (1000, 700, 896, 627), (500, 1333, 336, 896),
                         (448, 448, 448, 448), (300, 448, 448, 448), (31, 17, 5, 3), (2, 2, 448, 448), (2500, 1900, 1344, 1021)]:
        a = rs.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
        got = pre.resize(torch.from_numpy(a).cuda(), (w, h)).cpu().numpy()
        assert np.array_equal(got, np.asarray(Image.fromarray(a).resize((w, h)))), (W, H, w, h)
    # process_images: same crop count -> stacked, else list; PIL input accepted
    a = Image.fromarray(synthetic_image(0, 640, 480))
    out = pre.process_images([a, a])
    assert isinstance(out, torch.Tensor) and tuple(out.shape) == (2, 5, 3, 448, 448)
    out = pre.process_images([a, Image.fromarray(synthetic_image(1, 300, 900))])
    assert isinstance(out, list) and out[1].shape[0] == 4
    with pytest.raises(ValueError):
        pre(np.zeros((4, 4), dtype=np.uint8))


@pytest.mark.gpu
def test_gpu_hf_image_processor_surface():
    """OmChatImageProcessor (HF twin, image_processing_omchat.py:569-733): pixel_values zero-padded along the patch axis +
    num_patches; OmChatProcessor wires crops and placeholders together."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from omchat_b200.processing import OmChatImageProcessor, OmChatProcessor
    from toy_tokenizer import ToyTokenizer
    ip = OmChatImageProcessor(image_grid_pinpoints=PINPOINTS)
    a, b = synthetic_image(0, 640, 480), synthetic_image(1, 300, 900)
    out = ip([a, b])
    assert tuple(out.pixel_values.shape) == (2, 5, 3, 448, 448) and out.num_patches.tolist() == [5, 4]
    assert torch.equal(out.pixel_values[0].cpu(), torch.from_numpy(PO.process_anyres(a, PINPOINTS)))
    assert torch.equal(out.pixel_values[1, :4].cpu(), torch.from_numpy(PO.process_anyres(b, PINPOINTS)))
    assert float(out.pixel_values[1, 4].abs().max()) == 0.0
    feats = OmChatProcessor(ip, ToyTokenizer())("What is <image> this?", images=[a])
    assert tuple(feats.images.shape) == (5, 3, 448, 448) and feats.input_ids[0].tolist().count(-200) == 5
