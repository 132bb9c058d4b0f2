"""``from gnt.render_ray import render_rays`` (gnt/render_image.py, eval/gnt/*.py) -> nerfool_b200."""
from nerfool_b200.gnt.render_ray import render_rays  # noqa: F401
from nerfool_b200.render_ray import sample_pdf, sample_along_camera_ray  # noqa: F401
