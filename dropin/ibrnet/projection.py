"""``from ibrnet.projection import Projector`` (eval_adv.py:8, train.py) -> nerfool_b200."""
from nerfool_b200.projection import Projector  # noqa: F401
