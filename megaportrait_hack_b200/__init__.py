"""Import alias: the product package lives in `megaportrait-hack_b200/` (the directory name the build
contract fixes); a hyphen is not importable, so this stub makes `import megaportrait_hack_b200` resolve to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "megaportrait-hack_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
