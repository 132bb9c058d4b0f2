"""Drop-in overlay for the reference's ``gnt`` package (same mechanism as ``dropin/ibrnet``): ``gnt.projection``,
``gnt.transformer_network`` and ``gnt.render_ray`` resolve to nerfool_b200, every other ``gnt.*`` module to the
reference checkout named by NERFOOL_REFERENCE_ROOT."""
import os

_ref = os.environ.get('NERFOOL_REFERENCE_ROOT')
if _ref:
    _ref_pkg = os.path.join(_ref, 'gnt')
    if os.path.isdir(_ref_pkg) and _ref_pkg not in __path__:
        __path__.append(_ref_pkg)
