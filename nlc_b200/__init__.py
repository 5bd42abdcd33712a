"""Import shim: the package directory is named `diffusion-nlc_b200` (not a Python identifier), so
`import nlc_b200` maps onto it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "diffusion-nlc_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
