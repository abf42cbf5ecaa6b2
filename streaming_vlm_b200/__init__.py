"""Importable alias of the product package.

The product lives in the directory `streaming-vlm_b200/` (the name the build contract
asks for).  A hyphen is not a valid Python identifier, so this stub re-points the
package path at that directory and runs its `__init__.py`; `import streaming_vlm_b200`
and every `streaming_vlm_b200.<module>` resolve to files under `streaming-vlm_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "streaming-vlm_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
