"""Drop-in overlay for the reference's ``ibrnet`` package.

Put ``<repo>/dropin`` and ``<repo>`` ahead of the reference checkout on ``sys.path`` (or PYTHONPATH) and set
NERFOOL_REFERENCE_ROOT to the checkout: ``ibrnet.projection``, ``ibrnet.mlp_network`` and
``ibrnet.render_ray`` then resolve to the nerfool_b200 implementations while every other ``ibrnet.*`` module
(model.py, render_image.py, sample_ray.py, feature_network.py, criterion.py, data_loaders/ ...) is still the
reference's own file, unmodified (this package's ``__path__`` is extended with the reference's directory).
See INTEGRATION.md."""
import os

_ref = os.environ.get('NERFOOL_REFERENCE_ROOT')
if _ref:
    _ref_pkg = os.path.join(_ref, 'ibrnet')
    if os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
        __path__.append(_ref_pkg)
