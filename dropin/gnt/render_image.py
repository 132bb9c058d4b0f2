"""``from gnt.render_image import render_single_image`` (eval/gnt/eval.py, eval_adv.py, train.py) -> nerfool_b200."""
from nerfool_b200.gnt.render_image import render_single_image  # noqa: F401
