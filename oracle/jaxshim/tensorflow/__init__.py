"""Stand-in for the two tf.io.gfile calls nerfies/gpath.py makes (local files only).

TEST INFRASTRUCTURE: lets oracle/make_golden_camera.py import the reference's
unmodified nerfies/camera.py, whose only TensorFlow use is file IO."""
import os as _os
import types as _types

io = _types.SimpleNamespace(gfile=_types.SimpleNamespace(
    GFile=lambda path, *a, **k: open(str(path), *a, **k),
    exists=lambda p: _os.path.exists(str(p)),
    makedirs=lambda p: _os.makedirs(str(p), exist_ok=True),
    mkdir=lambda p: _os.mkdir(str(p)),
    glob=lambda p: __import__('glob').glob(str(p)),
    rmtree=lambda p: __import__('shutil').rmtree(str(p)),
    isdir=lambda p: _os.path.isdir(str(p)),
))
