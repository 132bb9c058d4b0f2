"""``from ibrnet.render_image import render_single_image`` (eval.py, eval_adv.py, train.py) -> nerfool_b200."""
from nerfool_b200.render_image import render_single_image  # noqa: F401
