"""Drop-in replacements for the reference model classes on the spectral-convolution path
(model/__init__.py:11-13 of the reference exports the same three names)."""
from .lanczos_net import *          # noqa: F401,F403
from .ada_lanczos_net import *      # noqa: F401,F403
from .lanczos_net_general import *  # noqa: F401,F403
from .gcn import *                  # noqa: F401,F403  (sibling models on the same kernels)
from .dcnn import *                 # noqa: F401,F403
from .cheby_net import *            # noqa: F401,F403
