"""``from gnt.projection import Projector`` -> nerfool_b200 (gnt/projection.py: the source cameras stay differentiable)."""
from nerfool_b200.gnt.projection import Projector  # noqa: F401
