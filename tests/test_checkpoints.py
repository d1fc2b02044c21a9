"""Flax checkpoint reader: the msgpack layout is restated from
flax 0.3's serialization.py (flax itself is not installable here: parity unpinned
against a file written by real flax).  The test vectors below are built by hand
from that layout, independently of the writer in nerfies_b200/checkpoints.py."""
import os

import msgpack
import numpy as np
import pytest
import torch

from nerfies_b200 import checkpoints as ck
from nerfies_b200 import model_utils


def _ext_array(a, code=1):
  a = np.asarray(a)
  return msgpack.ExtType(code, msgpack.packb((a.shape, a.dtype.name, a.tobytes()), use_bin_type=True))


def _hand_built_state(step=1234):
  rng = np.random.RandomState(0)
  params = {'model': {
      'warp_field': {'trunk': {'hidden_0': {'kernel': rng.randn(59, 128).astype(np.float32),
                                            'bias': rng.randn(128).astype(np.float32)}}},
      'nerf_mlps_coarse': {'MLP_0': {'hidden_0': {'kernel': rng.randn(51, 256).astype(np.float32)}}},
  }}
  enc = lambda t: {k: enc(v) for k, v in t.items()} if isinstance(t, dict) else _ext_array(t)
  tree = {'optimizer': {'target': enc(params),
                        'state': {'step': _ext_array(np.int32(step), 3), 'param_states': {}}},
          'warp_alpha': _ext_array(np.float32(6.5), 3), 'time_alpha': _ext_array(np.float32(0.25), 3)}
  return params, msgpack.packb(tree, use_bin_type=True)


def test_restore_hand_built_flax_bytes(tmp_path):
  params, blob = _hand_built_state(1234)
  (tmp_path / 'checkpoint_900').write_bytes(_hand_built_state(900)[1])
  (tmp_path / 'checkpoint_1234').write_bytes(blob)
  (tmp_path / 'checkpoint_1234.tmp').write_bytes(b'junk')
  assert ck.latest_checkpoint(str(tmp_path)).endswith('checkpoint_1234')      # 1234 > 900 numerically
  state = ck.restore_checkpoint(str(tmp_path))
  assert isinstance(state, model_utils.TrainState) and state.step == 1234
  assert state.warp_alpha == 6.5 and state.time_alpha == 0.25
  assert state.warp_extra == {'alpha': 6.5, 'time_alpha': 0.25}
  got = state.optimizer.target['model']['warp_field']['trunk']['hidden_0']['kernel']
  assert torch.is_tensor(got) and got.dtype == torch.float32 and got.shape == (59, 128)
  np.testing.assert_array_equal(got.numpy(), params['model']['warp_field']['trunk']['hidden_0']['kernel'])
  assert ck.restore_checkpoint(str(tmp_path), step=900).step == 900
  with pytest.raises(ValueError):
    ck.restore_checkpoint(str(tmp_path), step=5)


def test_restore_returns_target_when_directory_is_empty(tmp_path):
  sentinel = object()
  assert ck.restore_checkpoint(str(tmp_path), sentinel) is sentinel
  assert ck.restore_checkpoint(str(tmp_path / 'missing'), sentinel) is sentinel


def test_chunked_arrays_and_complex_leaves():
  a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
  tree = {'x': {'__msgpack_chunked_array__': True, 'shape': (2, 3, 4),
                'chunks': {'0': _ext_array(a.reshape(-1)[:10]), '1': _ext_array(a.reshape(-1)[10:])}},
          'c': msgpack.ExtType(2, msgpack.packb((1.5, -2.0))), 'i': 7}
  out = ck.msgpack_restore(msgpack.packb(tree, use_bin_type=True))
  np.testing.assert_array_equal(out['x'], a)
  assert out['c'] == complex(1.5, -2.0) and out['i'] == 7


def test_save_restore_round_trip_and_structure_check(tmp_path):
  from oracle import nerfies_oracle as O
  spec = O.OracleSpec(num_coarse_samples=8, num_fine_samples=8, use_warp=True, num_warp_embeddings=3,
                      num_appearance_embeddings=3, use_appearance_metadata=True)
  p = O.init_params(spec, 5)
  state = model_utils.TrainState(model_utils.Optimizer({'model': p}), warp_alpha=3.0, time_alpha=0.0)
  for step in (10, 20, 30):
    ck.save_checkpoint(str(tmp_path), state, step, keep=2)
  assert sorted(os.listdir(tmp_path)) == ['checkpoint_20', 'checkpoint_30']
  back = ck.restore_checkpoint(str(tmp_path), state)
  assert back.step == 30 and back.warp_alpha == 3.0

  def same(a, b):
    if isinstance(a, dict):
      assert set(a) == set(b)
      for k in a:
        same(a[k], b[k])
    else:
      assert torch.equal(torch.as_tensor(a), b)
  same(p, back.optimizer.target['model'])
  # structural check against a target with a different architecture
  other = O.init_params(O.OracleSpec(num_coarse_samples=8, num_fine_samples=8, use_warp=False), 5)
  with pytest.raises(ValueError):
    ck.restore_checkpoint(str(tmp_path), model_utils.TrainState(model_utils.Optimizer({'model': other})))
  with pytest.raises(ValueError):
    (tmp_path / 'checkpoint_99').write_bytes(msgpack.packb({'not': 'a state'}))
    ck.restore_checkpoint(str(tmp_path))
