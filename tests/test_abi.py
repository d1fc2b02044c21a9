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
