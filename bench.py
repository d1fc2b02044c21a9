"""Benchmark of the render hot path (NerfModel.__call__) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision P]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's CPU path (oracle port)

Metric (BASELINE.json): ray-samples/sec, coarse+fine, device-timed.  One
ray-sample = one (warp MLP + NeRF MLP) point evaluation; a ray costs
Nc + (Nc + Nf) of them.  A step = one forward of the whole
pipeline over one batch of synthetic rays.  Workload: the north-star synthetic
(65,536 rays x (128+128) samples, gpu_quarterhd.gin model dimensions) per GPU;
rays shard across GPUs with no data-path collective (weak scaling).

One JSON line is printed by rank 0 (see the task contract for the keys).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
  sys.path.insert(0, REPO)

FLOP_PER_RAY_SAMPLE = 1370112   # 2 x (97,792 + 587,264) MAC
NC, NF = 128, 128
EVALS_PER_RAY = NC + NC + NF
NEAR, FAR = 0.02, 0.83
N_IDS = 200


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--rays', type=int, default=65536, help='rays per GPU per step')
  ap.add_argument('--precision', default=None,
                  choices=[None, 'fp32', 'bf16', 'bf16x3'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--cpu-seconds', type=float, default=15.0)
  return ap.parse_args()


def oracle_spec():
  from oracle import nerfies_oracle as O
  return O.OracleSpec(
      num_coarse_samples=NC, num_fine_samples=NF, near=NEAR, far=FAR,
      num_nerf_point_freqs=8, sigma_activation='softplus', use_warp=True,
      warp_field_type='se3', use_appearance_metadata=True,
      num_warp_embeddings=N_IDS, num_appearance_embeddings=N_IDS)


def model_config():
  import nerfies_b200 as nb
  # gpu_quarterhd.gin (+ warp_defaults.gin, defaults.gin) model fields; the
  # deterministic path (eval.py:239).
  return nb.configs.ModelConfig(
      use_stratified_sampling=False, use_viewdirs=True, use_warp=True,
      warp_field_type='se3', num_warp_freqs=8, num_warp_features=8,
      use_appearance_metadata=True, sigma_activation='softplus',
      num_nerf_point_freqs=8, nerf_trunk_width=256, nerf_trunk_depth=8,
      num_coarse_samples=NC, num_fine_samples=NF)


def synthetic_rays(num_rays, seed):
  """SURVEY.md §8(d) synthetic inputs, float32, seeded."""
  import torch
  g = torch.Generator().manual_seed(seed)
  origins = torch.rand(num_rays, 3, generator=g) - 0.5
  d = torch.randn(num_rays, 3, generator=g)
  directions = d / torch.linalg.norm(d, dim=-1, keepdim=True)
  md = {'warp': torch.randint(0, N_IDS, (num_rays, 1), generator=g,
                              dtype=torch.int32),
        'appearance': torch.randint(0, N_IDS, (num_rays, 1), generator=g,
                                    dtype=torch.int32)}
  return {'origins': origins, 'directions': directions, 'metadata': md}


class ClockSampler:
  """nvidia-smi clocks / throttle reasons during the timed region."""
  FIELDS = ('clocks.sm,clocks.max.sm,power.draw,'
            'clocks_event_reasons.hw_slowdown,'
            'clocks_event_reasons.hw_thermal_slowdown,'
            'clocks_event_reasons.sw_thermal_slowdown,'
            'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}',
           '--format=csv,noheader,nounits', '-lms', '100'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if not self.proc:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, reasons = [], None, set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap']
    for r in self.rows:
      if len(r) < 7:
        continue
      try:
        sm.append(float(r[0]))
        mx = float(r[1])
      except ValueError:
        continue
      for n, v in zip(names, r[3:7]):
        if v.lower().startswith('active'):
          reasons.add(n)
    busy = [v for v in sm if mx and v > 0.3 * mx] or sm
    return {'sm_mhz': statistics.median(busy) if busy else None,
            'sm_max_mhz': mx, 'reasons': sorted(reasons),
            'samples': len(sm)}


def time_oracle(num_rays, threads, seed=0):
  """Seconds for one oracle forward of `num_rays` rays (torch CPU fp32)."""
  import torch
  from oracle import nerfies_oracle as O
  spec = oracle_spec()
  torch.set_num_threads(threads)
  params = O.make_trained_like(O.init_params(spec, seed), seed=seed + 1)
  rays = O.synthetic_rays(num_rays, spec, seed=seed + 2)
  chunk = 256   # bounds the (B, Nc-1, Nf) mask of sample_pdf and activations
  t0 = time.perf_counter()
  with torch.no_grad():
    for s in range(0, num_rays, chunk):
      sub = {'origins': rays['origins'][s:s + chunk],
             'directions': rays['directions'][s:s + chunk],
             'metadata': {k: v[s:s + chunk] for k, v in rays['metadata'].items()}}
      O.render_forward(params, spec, sub, warp_alpha=8.0)
  return time.perf_counter() - t0


def best_thread_count():
  """torch's intra-op pool does not scale to every core on large hosts (128
  threads on the GPU box is ~30x slower than 8-32): probe a few counts on a small
  sample and keep the fastest.  Returns (threads, rays_per_second)."""
  cores = os.cpu_count() or 1
  cands = sorted({c for c in (cores, 64, 32, 16, 8) if c <= cores})
  time_oracle(128, cands[0])                     # warm-up (thread pools, MKL)
  best = None
  for c in cands:
    t = time_oracle(128, c)
    if best is None or t < best[1]:
      best = (c, t)
  return best[0], 128 / best[1]


def cpu_baseline(budget_s):
  """The reference's algorithm on the host cores (oracle port; the JAX original
  cannot run in this image), on a bounded sample of the same workload."""
  threads, rate = best_thread_count()
  n = int(min(16384, max(256, rate * budget_s)) // 256 * 256)
  t = time_oracle(n, threads)
  return {'value': n * EVALS_PER_RAY / t, 'unit': 'ray-samples/s',
          'cores': threads, 'kind': 'port',
          'sample': f'{n} rays x ({NC}+{NF}) samples, quarterhd dims, '
                    f'torch-CPU fp32 oracle, {t:.1f} s; {threads} of '
                    f'{os.cpu_count()} host threads (fastest of a probe)'}


def run_reference(args):
  """--impl reference: the reference's CPU path (oracle port), all host threads,
  each step a bounded sample of the workload."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  threads, rate = best_thread_count()
  n = int(min(8192, max(256, rate * 8.0)) // 256 * 256)   # ~8 s per step
  for _ in range(min(args.warmup, 1)):
    time_oracle(n, threads)
  times = [time_oracle(n, threads) for _ in range(args.steps)]
  sec = sum(times) / len(times)
  value = n * EVALS_PER_RAY / sec
  line = {
      # same metric / unit / workload as the b200 arm (host-timed: there is no device)
      'impl': 'reference', 'metric': 'ray-samples/sec (coarse+fine, device-timed)',
      'value': value, 'unit': 'ray-samples/s', 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': min(args.warmup, 1),
      'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
      'config': {'workload': f'north-star synthetic: 65536 rays/GPU x ({NC}+{NF}) samples = '
                             f'{EVALS_PER_RAY} ray-samples/ray, gpu_quarterhd.gin model dims, SE(3) '
                             f'warp on, deterministic sampling, trained-like random weights; each '
                             f'step a bounded sample of {n} rays of it on the host CPU',
                 'rays_per_step': n, 'precision': 'fp32',
                 'timing': 'host wall clock around the reference algorithm (oracle port)'},
      'cpu_baseline': {'value': value, 'unit': 'ray-samples/s',
                       'cores': threads, 'kind': 'port',
                       'sample': f'{n} rays per step; restated reference on '
                                 'torch-CPU fp32 (JAX/Flax not installable)'},
      'e2e': {'value': value, 'unit': 'ray-samples/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  emit(line)


def run_b200(args):
  import torch
  import torch.distributed as dist
  import nerfies_b200 as nb

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  if args.gpus != world and rank == 0 and world > 1:
    print(f'warning: --gpus {args.gpus} but WORLD_SIZE={world}', file=sys.stderr)
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)

  precision = args.precision or default_precision()
  B = args.rays
  model, params = nb.construct_nerf(0, model_config(), B, range(N_IDS), [0],
                                    range(N_IDS), NEAR, FAR,
                                    precision=precision, device=dev)
  # "trained-like" weights: non-degenerate densities so the
  # resampled PDF and the composite do real work.
  from oracle import nerfies_oracle as O  # weights recipe only (not timed)
  cpu = lambda t: ({k: cpu(v) for k, v in t.items()} if isinstance(t, dict)
                   else t.cpu())
  gpu = lambda t: ({k: gpu(v) for k, v in t.items()} if isinstance(t, dict)
                   else t.to(dev))
  params = gpu(O.make_trained_like(cpu(params), seed=1))
  rays_host = synthetic_rays(B, seed=1000 + rank)   # each rank its own rays
  rays = {'origins': rays_host['origins'].to(dev),
          'directions': rays_host['directions'].to(dev),
          'metadata': {k: v.to(dev) for k, v in rays_host['metadata'].items()}}
  variables = {'params': params}
  warp_extra = {'alpha': 8.0, 'time_alpha': 0.0}
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

  def step():
    return model.apply(variables, rays, warp_extra=warp_extra)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(args.warmup):
    out = step()
  hd = model.handle(B)
  _ = hd.lib.nfb_set_profiling(hd.h, 1)
  barrier()
  sampler = ClockSampler(local_rank)
  sampler.start()
  launches0 = model.kernel_launches()
  ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for _ in range(args.steps)]
  field_ms = []
  barrier()
  wall0 = time.perf_counter()
  for i in range(args.steps):
    flush.zero_()                      # L2 flush between timed iterations
    ev[i][0].record()
    out = step()
    ev[i][1].record()
    ev[i][1].synchronize()
    field_ms.append((float(hd.lib.nfb_field_time_ms(hd.h, 0)),
                     float(hd.lib.nfb_field_time_ms(hd.h, 1))))
  barrier()
  wall = time.perf_counter() - wall0
  launches = model.kernel_launches() - launches0
  clocks = sampler.stop()
  hd.lib.nfb_set_profiling(hd.h, 0)
  step_ms = [a.elapsed_time(b) for a, b in ev]
  total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
  ms_per_step = float(total_ms) / args.steps
  value = world * B * EVALS_PER_RAY / (ms_per_step * 1e-3)

  # End to end through the C ABI's host entry point: host buffers in, host
  # buffers out, H2D + D2H inside the timed region.
  host_rays = {'origins': rays_host['origins'].numpy(),
               'directions': rays_host['directions'].numpy(),
               'metadata': {k: v.numpy() for k, v in rays_host['metadata'].items()}}
  model.apply_host(variables, host_rays, warp_extra=warp_extra)   # warm-up
  e2e_steps = max(2, min(args.steps, 5))
  barrier()
  t0 = time.perf_counter()
  for _ in range(e2e_steps):
    host_out = model.apply_host(variables, host_rays, warp_extra=warp_extra)
  barrier()
  e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev,
                       dtype=torch.float64)
  if world > 1:
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
  e2e_value = world * B * EVALS_PER_RAY / float(e2e_s)
  h2d = B * (12 + 12 + 4 + 4)          # origins, directions, warp id, appearance id
  d2h = B * 2 * 6 * 4                  # (B,6) per level

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  peaks = {}
  try:
    with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as f:
      peaks = json.load(f)
  except OSError:
    pass
  # The field kernel runs for ~all of a multi-hundred-ms step: sustained peak.
  peak = peaks.get('bf16_tflops_sustained') or peaks.get('bf16_tflops')
  peak_src = 'measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)'
  if not peak:
    peak, peak_src = 1590.0, 'fallback (B200_PROFILING.md)'
  fc = statistics.mean(m[0] for m in field_ms)
  ff = statistics.mean(m[1] for m in field_ms)
  # dominant kernel = the fine-level field launch (2/3 of the ray-samples).
  fine_flop = B * (NC + NF) * FLOP_PER_RAY_SAMPLE
  achieved = fine_flop / (ff * 1e-3) / 1e12
  traffic = None
  try:
    with open(os.path.join(REPO, 'profiles', 'traffic.json')) as f:
      traffic = json.load(f).get(precision, {}).get('field_fine_dram_bytes')
  except (OSError, ValueError):
    pass
  line = {
      'metric': 'ray-samples/sec (coarse+fine, device-timed)',
      'value': value, 'unit': 'ray-samples/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': {'fp32': 'fp32', 'bf16': 'bf16', 'bf16x3': 'bf16x3 (fp32-emulating)'}[precision],
      'data': 'synthetic',
      'config': {
          'workload': f'north-star synthetic: {B} rays/GPU x ({NC}+{NF}) samples '
                      f'= {EVALS_PER_RAY} ray-samples/ray, gpu_quarterhd.gin model '
                      'dims, SE(3) warp on, deterministic sampling, trained-like '
                      'random weights',
          'rays_per_gpu': B, 'precision': precision,
          'parallelism': f'ray sharding x{world}, no data-path collective',
          'l2': 'L2 flushed (256 MiB memset) between timed iterations; the '
                'per-step working set (268 MB of per-sample outputs) also '
                'exceeds L2',
          'timing': 'per-step CUDA events on the launch stream, summed, max '
                    'over ranks',
      },
      'clocks': clocks,
      'e2e': {'value': e2e_value, 'unit': 'ray-samples/s',
              'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
              'api': 'NerfModel.apply_host -> nfb_render_forward_host (host '
                     'buffers, pinned staging, H2D+D2H timed)'},
      'gpu_launches': int(launches),
      'roofline': {
          'bound': 'tensor', 'achieved': achieved, 'peak': peak,
          'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
          'kernel': 'field kernel, fine level '
                    f'({B * (NC + NF)} rows x {FLOP_PER_RAY_SAMPLE} FLOP)',
          'kernel_ms': ff, 'coarse_kernel_ms': fc,
          'share_of_step': (fc + ff) / ms_per_step,
          'peak_source': peak_src,
          'note': ('fp32 parity mode runs on the FFMA pipe; the fraction is '
                   'still quoted against the bf16 tensor peak the north-star names'
                   if precision == 'fp32' else ''),
      },
      'wall_s_timed_region': wall,
  }
  if not args.no_cpu_baseline:
    line['cpu_baseline'] = cpu_baseline(args.cpu_seconds)
  emit(line)
  if world > 1:
    dist.destroy_process_group()


def default_precision():
  """The fastest mode the built library contains."""
  return 'bf16' if os.path.exists(os.path.join(
      REPO, 'nerfies_b200', 'csrc', 'field_tc.cuh')) else 'fp32'


_SAVED_STDOUT = None


def quiet_stdout():
  """Route fd 1 to stderr until emit(): libraries (NCCL's version banner, ...)
  must not print in front of the one JSON line the driver parses."""
  global _SAVED_STDOUT
  sys.stdout.flush()
  _SAVED_STDOUT = os.dup(1)
  os.dup2(2, 1)


def emit(line):
  sys.stdout.flush()
  if _SAVED_STDOUT is not None:
    os.dup2(_SAVED_STDOUT, 1)
  print(json.dumps(line), flush=True)


def main():
  args = parse_args()
  quiet_stdout()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
