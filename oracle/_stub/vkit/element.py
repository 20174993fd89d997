"""Stand-in for the un-vendored ``vkit`` package, used ONLY by oracle/make_golden.py to import the reference in the build
container.  On the training path the reference touches nothing of ``vkit`` but the four inclusive bounds of ``Box``
(loss_function/adaptive_scaling.py:15,75-86); the tensor side of the inference passes (inferencing/adaptive_scaling.py:
92-188, 295-396) additionally needs records that carry a numpy ``mat`` (Image / Mask / ScoreMap).  The geometry classes are
import-only placeholders: nothing the golden generator runs calls them."""
import attrs
import numpy as np


@attrs.define
class Box:
    up: int
    down: int
    left: int
    right: int


@attrs.define
class Image:
    mat: np.ndarray

    @property
    def height(self) -> int:
        return int(self.mat.shape[0])

    @property
    def width(self) -> int:
        return int(self.mat.shape[1])

    def to_rgb_image(self):
        assert self.mat.ndim == 3 and self.mat.shape[2] == 3
        return self


@attrs.define
class Mask:
    mat: np.ndarray

    @property
    def np_mask(self):
        return self.mat.astype(bool)


@attrs.define
class ScoreMap:
    mat: np.ndarray
    is_prob: bool = True


class Point:        # placeholders (never instantiated by the golden generator)
    pass


class PointTuple(tuple):
    pass


class Polygon:
    pass
