"""Import alias: the package directory is `tg-pose_b200/` (not a valid Python identifier), so
`import tgpose_b200` resolves its submodules there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tg-pose_b200")]
exec(open(_os.path.join(__path__[0], "__init__.py")).read())
