"""Import-only stand-in (see vkit/element.py)."""
