"""Optional on-screen viewer for `render_mode="human"`.

The reference opens a pygame window from inside the env class (`src/rbc_gym/envs/rbc2D.py:214-261`).  The window is not part of
the accelerated path; this module only puts the RGB frame the env already produced (`render()` of the env classes,
`rbc2d_render_rgb_dev` for whole batches) on the screen, and needs pygame only when it is actually used."""
from __future__ import annotations

import numpy as np


class FrameWindow:
    """A fixed-size window that shows `(width, height, 3)` uint8 frames, rate-limited to `fps`."""

    def __init__(self, width: int, height: int, fps: float, title: str = "Rayleigh Benard Convection"):
        try:
            import pygame
        except ImportError as e:
            raise RuntimeError("render_mode='human' needs pygame, which is not installed; use render_mode='rgb_array'") from e
        self._pg = pygame
        pygame.display.init()
        self._size = (int(width), int(height))
        self._window = pygame.display.set_mode(self._size)
        pygame.display.set_caption(title)
        self._tick = pygame.time.Clock()
        self._fps = fps

    def show(self, frame_xy: np.ndarray) -> None:
        """`frame_xy[x, y]` = RGB of pixel column x, row y (pygame's surface-array convention)."""
        pg = self._pg
        picture = pg.surfarray.make_surface(np.ascontiguousarray(frame_xy, dtype=np.uint8))
        pg.transform.scale(picture, self._size, self._window)
        pg.event.pump()
        pg.display.flip()
        self._tick.tick(self._fps)

    def close(self) -> None:
        self._pg.display.quit()
        self._pg.quit()
