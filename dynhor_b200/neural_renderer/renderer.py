"""`nr.renderer.Renderer` (utils/losses.py:36)."""
from ..renderer import Renderer  # noqa: F401
