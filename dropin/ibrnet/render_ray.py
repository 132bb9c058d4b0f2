"""``from ibrnet.render_ray import render_rays`` (eval_adv.py, train.py, render_image.py:18) -> nerfool_b200."""
from nerfool_b200.render_ray import (render_rays, render_rays_hybrid, sample_pdf, raw2outputs,  # noqa: F401
                                     sample_along_camera_ray)
