"""Deterministic synthetic RGB test images (numpy legacy RandomState: bit-stable across numpy versions), shared by the golden
generator and the tests so that the fixture only has to store outputs."""
import numpy as np

PINPOINTS = [[448, 896], [896, 448], [896, 896], [1344, 448], [448, 1344], [1344, 1344]]
SIZES = [(640, 480), (300, 900), (1500, 500), (448, 448), (97, 131), (1300, 1250), (1001, 333)]  # (W, H)


def synthetic_image(n: int, W: int, H: int) -> np.ndarray:
    rs = np.random.RandomState(1000 + n)
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.empty((H, W, 3), dtype=np.float64)
    for c in range(3):
        fx, fy, ph = rs.uniform(0.01, 0.2), rs.uniform(0.01, 0.2), rs.uniform(0, 6.28)
        img[:, :, c] = 127.5 + 90.0 * np.sin(fx * x + fy * y + ph) + 30.0 * np.cos(0.5 * fx * x - 0.7 * fy * y)
    img += rs.randint(-25, 26, size=img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
