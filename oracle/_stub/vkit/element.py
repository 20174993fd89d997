"""Stand-in for the un-vendored ``vkit`` package, used ONLY by oracle/make_golden.py to import the reference's
loss functions in the build container.  The reference touches nothing of ``vkit`` on this path but the four
inclusive bounds of ``Box`` (loss_function/adaptive_scaling.py:15,75-86)."""
import attrs


@attrs.define
class Box:
    up: int
    down: int
    left: int
    right: int
