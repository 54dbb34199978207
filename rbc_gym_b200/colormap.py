"""Turbo colormap without matplotlib (the reference renders with `matplotlib.colormaps["turbo"]`,
`src/rbc_gym/envs/rbc2D.py:23-26`).  7-term polynomial fit of turbo published by Google (A. Mikhailov,
Apache-2.0); max abs error ~1/255 per channel, rendering only."""
import numpy as np

_R = (0.13572138, 4.61539260, -42.66032258, 132.13108234, -152.94239396, 59.28637943)
_G = (0.09140261, 2.19418839, 4.84296658, -14.18503333, 4.27729857, 2.82956604)
_B = (0.10667330, 12.64194608, -60.58204836, 110.36276771, -89.90310912, 27.34824973)


def turbo(value, vmin=1.0, vmax=2.0) -> np.ndarray:
    """value[...] -> uint8 RGB[..., 3]."""
    x = np.clip((np.asarray(value, dtype=np.float64) - vmin) / (vmax - vmin), 0.0, 1.0)
    out = []
    for c in (_R, _G, _B):
        acc = np.zeros_like(x)
        for k in reversed(c):
            acc = acc * x + k
        out.append(acc)
    return (np.clip(np.stack(out, axis=-1), 0.0, 1.0) * 255).astype(np.uint8)
