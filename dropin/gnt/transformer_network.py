"""``from gnt.transformer_network import GNT`` (gnt/model.py:5) -> nerfool_b200."""
from nerfool_b200.gnt.transformer_network import GNT  # noqa: F401
