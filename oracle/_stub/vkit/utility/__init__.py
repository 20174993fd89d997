"""Import-only stand-in (see vkit/element.py)."""
from typing import Any
PathType = Any

