"""``from ibrnet.mlp_network import IBRNet`` (ibrnet/model.py:18) -> nerfool_b200."""
from nerfool_b200.mlp_network import IBRNet  # noqa: F401
