"""GNT path of the reference (``gnt/``; SURVEY.md 8 row a15 / f3, BASELINE config 5): drop-in ``GNT`` module and
``render_rays`` whose forward runs the CUDA kernels of ``csrc/nfb_gnt.cu``.  Forward (render) only in this round."""
from .transformer_network import GNT          # noqa: F401
from .render_ray import render_rays           # noqa: F401
