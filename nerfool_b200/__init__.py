"""nerfool_b200: B200-native (sm_100a) per-ray generalizable-NeRF hot path behind the reference's
``Projector.compute`` / ``IBRNet.forward`` / ``render_rays`` API.  See DESIGN.md."""
from .projection import Projector          # noqa: F401
from .mlp_network import IBRNet            # noqa: F401
from .render_ray import render_rays, sample_pdf, raw2outputs, sample_along_camera_ray  # noqa: F401
from ._lib import set_precision, get_precision  # noqa: F401

__all__ = ['Projector', 'IBRNet', 'render_rays', 'sample_pdf', 'raw2outputs', 'sample_along_camera_ray',
           'set_precision', 'get_precision']
