"""``from gnt.projection import Projector`` -> nerfool_b200 (gnt/projection.py equals ibrnet/projection.py up to formatting)."""
from nerfool_b200.projection import Projector  # noqa: F401
