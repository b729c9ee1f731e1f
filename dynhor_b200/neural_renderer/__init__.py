"""`import neural_renderer as nr` shim: put dynhor_b200/ on sys.path ahead of site-packages (or alias
sys.modules['neural_renderer'] = dynhor_b200.neural_renderer) and the reference's utils/losses.py:5,36-40,68
and pose_initializtion.py:29,98-105 run on the B200 kernels unmodified.  Only the silhouette / projection
subset the reference calls is provided."""
from ..renderer import Renderer, projection  # noqa: F401
from . import renderer  # noqa: F401
