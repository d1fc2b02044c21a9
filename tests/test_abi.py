"""The C-ABI library loads without a GPU and exports every symbol the header
declares; with no device, handle creation fails loudly (no CPU fallback)."""
import ctypes
import os
import re

from nerfies_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
  text = open(os.path.join(REPO, 'include', 'nerfies_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(nfb_[a-z_0-9]+)\s*\(', text)))


def test_header_and_binding_agree():
  assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
  lib = _lib.load()
  for name in _header_symbols():
    assert hasattr(lib, name), f'{name} not exported'
  assert lib.nfb_version().decode().startswith('nerfies_b200')


def test_config_struct_matches_header():
  text = open(os.path.join(REPO, 'include', 'nerfies_b200.h')).read()
  body = re.search(r'typedef struct nfb_config \{(.*?)\} nfb_config;', text,
                   re.S).group(1)
  body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
  names = []
  for decl in body.split(';'):
    decl = decl.strip()
    if not decl:
      continue
    decl = re.sub(r'^(unsigned|int|float)\s+', '', decl)
    names += [n.strip() for n in decl.split(',')]
  assert names == [f[0] for f in _lib.NfbConfig._fields_]


def test_no_device_is_a_loud_error():
  import torch
  if torch.cuda.is_available():
    return
  lib = _lib.load()
  cfg = _lib.NfbConfig()
  cfg.num_coarse_samples = 8
  h = ctypes.c_void_p()
  assert lib.nfb_create(ctypes.byref(cfg), 4, ctypes.byref(h)) != 0
  assert b'no CUDA device' in lib.nfb_last_error()


def test_camera_struct_matches_header():
  text = open(os.path.join(REPO, 'include', 'nerfies_b200.h')).read()
  body = re.search(r'typedef struct nfb_camera \{(.*?)\} nfb_camera;', text, re.S).group(1)
  body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
  fields = []
  for decl in body.split(';'):
    decl = decl.strip()
    if not decl:
      continue
    m = re.match(r'(int|float)\s+([a-z_]+)(?:\[(\d+)\])?$', decl)
    assert m, decl
    fields.append((m.group(2), m.group(1), int(m.group(3) or 1)))
  got = []
  for name, ctype in _lib.NfbCamera._fields_:
    n = getattr(ctype, '_length_', 1)
    base = ctype._type_ if n > 1 or hasattr(ctype, '_length_') else ctype
    got.append((name, 'float' if base is ctypes.c_float else 'int', n))
  assert fields == got
  assert ctypes.sizeof(_lib.NfbCamera) == 4 * sum(n for _, _, n in fields)


def test_camera_entry_points_fail_loudly_without_a_device():
  import torch
  if torch.cuda.is_available():
    return
  lib = _lib.load()
  cam = _lib.NfbCamera()
  cam.focal_length = 100.0
  cam.pixel_aspect_ratio = 1.0
  cam.image_size[:] = [4, 4]
  assert lib.nfb_camera_rays(ctypes.byref(cam), 0, 16, None, ctypes.c_void_p(16), None, None) != 0
  assert b'no CUDA device' in lib.nfb_last_error()
  assert lib.nfb_camera_rays(ctypes.byref(cam), 10, 16, None, ctypes.c_void_p(16), None, None) != 0
  assert b'exceeds' in lib.nfb_last_error()
