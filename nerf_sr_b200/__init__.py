"""nerf_sr_b200 -- B200-native (sm_100a) implementation of NeRF-SR's volumetric-render hot path.

The product is ``libnsr_b200.so`` (C ABI in include/nsr.h); this package is the thin Python host
layer that mirrors the reference's ``forward_rays`` / ``render_rays`` interface.  Importing the
package does not load the library; constructing a ``Renderer`` does, and fails loudly without it."""
from ._lib import NsrError, PRECISIONS, LIB_PATH  # noqa: F401
from .renderer import Renderer, config_from_opt, patch_model, state_dict_order  # noqa: F401
from .training import RenderFunction, Trainer  # noqa: F401

__all__ = ["Renderer", "Trainer", "RenderFunction", "NsrError", "config_from_opt", "patch_model", "state_dict_order",
           "PRECISIONS", "LIB_PATH"]
