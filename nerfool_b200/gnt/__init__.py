"""GNT path of the reference (``gnt/``; SURVEY.md 8 row a15 / f3, BASELINE config 5): drop-in ``GNT`` module and
``render_rays`` whose forward runs the CUDA kernels of ``csrc/nfb_gnt.cu`` and whose data gradient (``csrc/nfb_gnt_bwd.cu``) carries the attack
of eval/gnt/eval_adv.py back to the feature maps and the source cameras."""
from .transformer_network import GNT          # noqa: F401
from .render_ray import render_rays           # noqa: F401
from .projection import Projector            # noqa: F401
