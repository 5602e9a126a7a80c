"""Import alias: the product package lives in ``rcnn-ocr_b200/`` (a directory name Python
cannot import directly); this shim makes it importable as ``rcnn_ocr_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rcnn-ocr_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r", encoding="utf-8") as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
