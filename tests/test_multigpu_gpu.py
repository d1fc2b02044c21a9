"""Frame rendering on 1 and 2 GPUs with the REAL kernels:
an N-GPU sharded render equals the 1-GPU render bit for bit.

eval.py:330-353 + evaluation.py:52-95: a frame's rays are split over the devices,
every device renders its shard, one all_gather assembles the frame.  Rays are
independent (no cross-ray operation on the path), so the split cannot change a bit.
The 2-rank tests need two GPUs (`gpurun --gpus 2`); on a 1-GPU box they are skipped.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _camera(w=96, h=54):
  import nerfies_b200 as nb
  th = 0.3
  R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], np.float32)
  return nb.camera.Camera(orientation=R, position=[0.1, -0.05, -0.4], focal_length=80.0,
                          principal_point=[w / 2, h / 2], image_size=[w, h],
                          radial_distortion=[0.02, -0.01, 0.0], tangential_distortion=[1e-3, -5e-4])


def _model(device, precision):
  import nerfies_b200 as nb
  from oracle import nerfies_oracle as O
  cfg = nb.configs.ModelConfig(
      use_stratified_sampling=False, use_warp=True, warp_field_type='se3',
      use_appearance_metadata=True, num_coarse_samples=64, num_fine_samples=64,
      num_nerf_point_freqs=8, sigma_activation='softplus')
  model, params = nb.construct_nerf(0, cfg, 2048, range(10), [0], range(10), near=0.02,
                                    far=0.83, precision=precision, device=device)
  cpu = lambda t: ({k: cpu(v) for k, v in t.items()} if isinstance(t, dict) else t.cpu())
  dev = lambda t: ({k: dev(v) for k, v in t.items()} if isinstance(t, dict) else t.to(device))
  return model, dev(O.make_trained_like(cpu(params), seed=4))


@pytest.mark.parametrize('precision', ['fp16x3', 'bf16'])
def test_render_frame_equals_render_image_on_one_gpu(precision):
  """render_frame (rays generated on the GPU per slab, one launch per slab) ==
  render_image over the same frame's rays in reference-sized chunks."""
  from nerfies_b200 import evaluation, camera as camera_lib
  from nerfies_b200.model_utils import Optimizer, TrainState
  dev = torch.device('cuda', 0)
  model, params = _model(dev, precision)
  cam = _camera()
  md = {'warp': 3, 'appearance': 7}
  extra = {'alpha': 6.5, 'time_alpha': 0.0}
  frame = evaluation.render_frame(model, params, cam, extra, md, max_rays=2048)
  rays = camera_lib.camera_to_rays(cam, dev)
  h, w = cam.image_shape
  rays_dict = {'origins': rays['origins'], 'directions': rays['directions'],
               'metadata': {k: torch.full((h, w, 1), v, dtype=torch.int32, device=dev) for k, v in md.items()}}
  state = TrainState(Optimizer({'model': params}), warp_alpha=6.5)
  ref = evaluation.render_image(state, rays_dict, evaluation.make_model_fn(model), 1, 0, chunk=1000)
  torch.cuda.synchronize()
  for k in ('rgb', 'depth', 'med_depth', 'acc'):
    assert frame[k].shape == ref[k].shape
    assert torch.equal(frame[k], ref[k]), k


def _worker(rank, world, port, tmp, precision):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                    WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  sys.path.insert(0, ROOT)
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
  try:
    from nerfies_b200 import evaluation, camera as camera_lib
    from nerfies_b200.model_utils import Optimizer, TrainState
    dev = torch.device('cuda', rank)
    model, params = _model(dev, precision)
    cam = _camera(97, 53)                   # 5141 rays: not a multiple of 2 -> edge padding
    md = {'warp': 3, 'appearance': 7}
    extra = {'alpha': 6.5, 'time_alpha': 0.0}
    t = {}
    frame = evaluation.render_frame(model, params, cam, extra, md, max_rays=1500, timings=t)
    # the reference-shaped path: render_image with device_count = world
    rays = camera_lib.camera_to_rays(cam, dev)
    h, w = cam.image_shape
    rays_dict = {'origins': rays['origins'], 'directions': rays['directions'],
                 'metadata': {k: torch.full((h, w, 1), v, dtype=torch.int32, device=dev)
                              for k, v in md.items()}}
    state = TrainState(Optimizer({'model': params}), warp_alpha=6.5)
    img = evaluation.render_image(state, rays_dict, evaluation.make_model_fn(model), world, 0, chunk=999)
    torch.cuda.synchronize()
    torch.save({'frame': {k: v.cpu() for k, v in frame.items()},
                'image': {k: v.cpu() for k, v in img.items()}, 'timings': t},
               os.path.join(tmp, f'out{rank}.pt'))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('precision', ['fp16x3'])
def test_two_gpu_frame_is_bit_identical_to_one_gpu(tmp_path, precision):
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
  from nerfies_b200 import evaluation
  port = 29700 + os.getpid() % 1500
  mp.spawn(_worker, args=(2, port, str(tmp_path), precision), nprocs=2, join=True)
  dev = torch.device('cuda', 0)
  model, params = _model(dev, precision)
  single = evaluation.render_frame(model, params, _camera(97, 53), {'alpha': 6.5, 'time_alpha': 0.0},
                                   {'warp': 3, 'appearance': 7}, max_rays=4096)
  torch.cuda.synchronize()
  for r in range(2):
    out = torch.load(os.path.join(str(tmp_path), f'out{r}.pt'))
    for k in ('rgb', 'depth', 'med_depth', 'acc'):
      assert torch.equal(out['frame'][k], single[k].cpu()), ('render_frame', r, k)
      assert torch.equal(out['image'][k], single[k].cpu()), ('render_image', r, k)
