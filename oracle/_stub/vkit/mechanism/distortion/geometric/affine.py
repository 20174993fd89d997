"""Import-only stand-in (see vkit/element.py)."""
affine_polygons = RotateConfig = RotateState = None

