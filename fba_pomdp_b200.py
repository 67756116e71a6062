"""Import shim: the package directory is named `fba-pomdp_b200/` (not a valid Python identifier),
so `import fba_pomdp_b200` resolves here and loads that directory as the package."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "fba-pomdp_b200")]
__package__ = __name__
__file__ = _os.path.join(__path__[0], "__init__.py")
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
