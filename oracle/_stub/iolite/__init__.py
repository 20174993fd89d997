"""Import-only stand-in (see vkit/element.py)."""
def folder(path, **kwargs):
    return path

