from . import lpips_tf  # noqa: F401
