#!/usr/bin/env python
"""bench.py -- QM8-shaped molecules/sec of the LanczosNet spectral-convolution forward.

    python bench.py --gpus N --steps K --warmup W            # this repo (B200, sm_100a)
    python bench.py --impl reference --gpus N ...            # CPU baseline (oracle port)

Workload (BASELINE.json configs[1]): LanczosNet forward, config/qm8_lanczos_net.yaml,
K=20 Ritz pairs, batch 1024 per GPU, synthetic QM8-shaped molecules (n_b in [3,26], N=26),
numpy-seeded weights, fp32 (the big Linear runs as 3xTF32 on tcgen05 = fp32-grade accuracy).
A step = one forward over one batch.  ``value`` = molecules/s with inputs resident in HBM;
``e2e`` = the same through the module's public forward() with pinned HOST inputs (H2D of
node_feat/L/D/V/mask and D2H of the scores inside the timed region).  Multi-GPU: one process
per GPU (torchrun), batch shards with no data-path collective, one NCCL all-gather of the
[B,16] predictions per step; weak scaling.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

BATCH = 1024
NUM_BATCHES = 8          # distinct resident batches rotated between steps (8 x 21 MB > L2 126 MB)
WEIGHT_SEED = 1234


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as fh:
      p = json.load(fh)
    return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
            'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
            'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
          'source': 'fallback'}


def make_batches(num, batch, seed0):
  from lanczosnetwork_b200 import data
  return [data.synthetic_qm8_batch(batch, seed=seed0 + i) for i in range(num)]


def build_model():
  from helpers import deterministic_state_dict
  from lanczosnetwork_b200 import configs
  from lanczosnetwork_b200.model import LanczosNet
  mod = LanczosNet(configs.qm8_lanczos_net())
  params = deterministic_state_dict(mod, WEIGHT_SEED)
  mod.load_state_dict(params)
  return mod, params


class ClockSampler(threading.Thread):
  """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

  def __init__(self, index):
    super(ClockSampler, self).__init__(daemon=True)
    self.index = index
    self.samples = []
    self.reasons = set()
    self.max_mhz = None
    self._halt = threading.Event()

  def run(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
      names = {
          pynvml.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
          pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
          pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
          pynvml.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
      }
      while not self._halt.is_set():
        self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for bit, nm in names.items():
          if r & bit:
            self.reasons.add(nm)
        time.sleep(0.002)
    except Exception as exc:   # NVML unavailable: report that instead of a number
      self.reasons.add('nvml_error:%s' % type(exc).__name__)

  def stop(self):
    self._halt.set()
    self.join(timeout=2.0)
    med = float(np.median(self.samples)) if self.samples else None
    return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
            'samples': len(self.samples)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_forward(params, spec, batch):
  from oracle import lanczos_oracle as orc
  return orc.lanczos_net_forward(params, spec, batch['node_feat'], batch['L'], batch['D'],
                                 batch['V'], batch['node_mask'])


def pick_cpu_threads(params, spec, batch):
  """The torch CPU path of the reference is made of thousands of tiny ops; on a many-core host
  all-cores is far from the best setting, so probe a few thread counts and keep the fastest."""
  cores = os.cpu_count() or 1
  small = {k: v[:64] for k, v in batch.items()}
  best, best_t = cores, None
  for nt in sorted(set([min(4, cores), min(8, cores), min(16, cores), min(32, cores), cores])):
    torch.set_num_threads(nt)
    cpu_reference_forward(params, spec, small)
    t0 = time.perf_counter()
    cpu_reference_forward(params, spec, small)
    dt = time.perf_counter() - t0
    if best_t is None or dt < best_t:
      best, best_t = nt, dt
  torch.set_num_threads(best)
  return best


def time_cpu_baseline(params, spec, batch, iters, warmup=1):
  for _ in range(warmup):
    cpu_reference_forward(params, spec, batch)
  ts = []
  for _ in range(iters):
    t0 = time.perf_counter()
    cpu_reference_forward(params, spec, batch)
    ts.append(time.perf_counter() - t0)
  return float(np.median(ts))


def run_reference_arm(args):
  """CPU baseline: the oracle port of the reference forward (the Python reference cannot
  travel to the GPU box), all host threads, one bounded-sample batch per step."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  from helpers import oracle_spec
  mod, params = build_model()
  spec = oracle_spec(mod, 'LanczosNet')
  sample = 256
  batch = make_batches(1, sample, 4242)[0]
  cores = pick_cpu_threads(params, spec, batch)
  for _ in range(max(args.warmup, 1)):
    cpu_reference_forward(params, spec, batch)
  t0 = time.perf_counter()
  for _ in range(args.steps):
    cpu_reference_forward(params, spec, batch)
  dt = time.perf_counter() - t0
  value = sample * args.steps / dt
  line = {
      'impl': 'reference', 'metric': 'QM8 molecules/sec (forward)', 'value': value,
      'unit': 'molecules/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'QM8 LanczosNet forward (config/qm8_lanczos_net.yaml), K=20, '
                             'N=26; CPU oracle port of model/lanczos_net.py on %d-molecule '
                             'batches' % sample},
      'cpu_baseline': {'value': value, 'unit': 'molecules/s', 'cores': cores, 'kind': 'port',
                       'sample': '%d steps x %d molecules, torch CPU fp32, best of {4,8,16,32,all} '
                                 'threads = %d (host has %d cores)'
                                 % (args.steps, sample, cores, os.cpu_count() or 1)},
      'e2e': {'value': value, 'unit': 'molecules/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--batch', type=int, default=BATCH)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

  if args.impl == 'reference':
    run_reference_arm(args)
    return

  import torch.distributed as dist
  from helpers import oracle_spec
  from lanczosnetwork_b200 import ops, sharded

  rank, world, local = sharded.init_from_env('nccl')
  if world != args.gpus:
    raise SystemExit('--gpus %d but WORLD_SIZE=%d (launch with torchrun)' % (args.gpus, world))
  dev = torch.device('cuda', local)
  torch.cuda.set_device(dev)

  mod, params = build_model()
  spec = oracle_spec(mod, 'LanczosNet')
  mod = mod.to(dev).eval()
  B = args.batch
  host = make_batches(NUM_BATCHES, B, 1000 + 100 * rank)      # per-rank shard (weak scaling)
  keys = ('node_feat', 'L', 'D', 'V', 'node_mask')
  pinned = [{k: torch.from_numpy(b[k]).pin_memory() for k in keys} for b in host]
  resident = [{k: v.to(dev) for k, v in p.items()} for p in pinned]
  h2d_bytes = sum(v.numel() * v.element_size() for v in pinned[0].values())
  P = 16
  gathered = torch.empty((B * world, P), device=dev) if world > 1 else None
  out_host = torch.empty((B * world if world > 1 else B, P)).pin_memory()

  def step_resident(i):
    b = resident[i % NUM_BATCHES]
    score = mod(b['node_feat'], b['L'], b['D'], b['V'], mask=b['node_mask'])
    if world > 1:
      dist.all_gather_into_tensor(gathered, score)
      return gathered
    return score

  def step_e2e(i):
    p = pinned[i % NUM_BATCHES]
    score = mod(p['node_feat'], p['L'], p['D'], p['V'], mask=p['node_mask'])   # H2D inside
    if world > 1:
      dist.all_gather_into_tensor(gathered, score)
      score = gathered
    out_host.copy_(score, non_blocking=True)                                   # D2H
    return score

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
      fn(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())

  with torch.no_grad():
    # correctness gate on this rank's first batch (small slice, CPU oracle as the checker)
    if rank == 0:
      chk = {k: host[0][k][:32] for k in keys}
      ref = cpu_reference_forward(params, spec, chk).numpy()
      got = mod(*[resident[0][k][:32] for k in ('node_feat', 'L', 'D', 'V')],
                mask=resident[0]['node_mask'][:32]).cpu().numpy()
      max_err = float(np.abs(got - ref).max())
      if not np.allclose(got, ref, rtol=1e-4, atol=2e-5):
        raise SystemExit('bench: CUDA forward disagrees with the oracle (max err %g)' % max_err)
    else:
      max_err = None

    # every resident batch is seen twice before timing: the second sighting of a set of device
    # buffers is when the module captures its zero-copy graph for them
    for i in range(max(args.warmup, 2 * NUM_BATCHES)):
      step_resident(i)
    for i in range(args.warmup):
      step_e2e(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = ops.launch_count() - l0
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()

    # dominant kernel: the whole 7-layer spectral-conv stack + readout as ONE persistent tcgen05
    # kernel (K-depth 960 + 6 x 1920 per row), timed per launch with CUDA events on the
    # launching stream.
    events = []
    orig = ops.spectral_stack_forward

    def probed(prep, Q, w_hi, w_lo, bias, dins, H, S, **kw):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      torch.cuda._sleep(400000)           # GPU spins ~0.2 ms: the host enqueues event, kernel,
      a.record()                          # event meanwhile, so no launch latency sits between them
      r = orig(prep, Q, w_hi, w_lo, bias, dins, H, S, **kw)
      b.record()
      E1 = prep[0].shape[1]
      kdim = sum((S + E1) * d for d in dins)          # summed GEMM depth of all layers
      events.append((a, b, Q.shape[0] * Q.shape[1], H, kdim))
      return r

    ops.spectral_stack_forward = probed
    mod.use_cuda_graph = False            # eager launches so the events bracket single kernels
    for i in range(min(args.steps, 5)):
      step_resident(i)
    torch.cuda.synchronize(dev)
    mod.use_cuda_graph = True
    ops.spectral_stack_forward = orig

  peaks = load_peaks()
  roof = None
  if events:
    durs = [a.elapsed_time(b) for a, b, _, _, _ in events]
    M, N, K = events[0][2], events[0][3], events[0][4]
    flops = 2.0 * M * N * K                       # algorithmic (padded B*N rows) flops per launch
    avg_ms = float(np.mean(durs))
    achieved = flops / (avg_ms * 1e-3) / 1e12
    peak_tf32 = peaks['bf16_tflops_sustained'] / 2.0
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'dominant_kernel_traffic.json')
    if os.path.exists(tpath):
      with open(tpath) as fh:
        traffic = json.load(fh).get('dram_bytes_per_launch')
    # what the tensor pipe really executes: packed 128-row tiles x 3 TF32 MMAs per product
    prep = ops.graph_prepare(resident[0]['L'], resident[0]['V'])
    n_tiles = int(prep[4][0].item())
    real_rows = int(prep[3][:, 0].sum().item())
    executed = 3.0 * 2.0 * n_tiles * 128 * N * K / (avg_ms * 1e-3) / 1e12
    roof = {
        'bound': 'tensor', 'kernel': 'tc_gemm_kernel<SpectralPolicy> (lnb_spectral_stack_forward, 7 layers + readout)',
        'achieved': achieved, 'peak': peak_tf32, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf32,
        'traffic': traffic, 'avg_ms_per_launch': avg_ms, 'launch_shape': [M, N, K],
        'executed_tensor_tflops': executed, 'frac_executed': executed / peak_tf32,
        'useful_tflops': 2.0 * real_rows * N * K / (avg_ms * 1e-3) / 1e12,
        'packed_tiles': n_tiles, 'real_rows': real_rows,
        'note': 'achieved = ALGORITHMIC fp32-equivalent GEMM flops 2*(B*N)*H*sum_l(C*D_l) of the padded '
                'reference formulation / CUDA-event time per launch. The kernel drops padded rows '
                '(packed tiles) and issues 3 TF32 MMAs per product (3xTF32): executed_tensor_tflops '
                '= 3*2*(tiles*128)*H*(C*D)/t is what the tensor pipe does; useful_tflops counts real '
                'nodes only. peak = %s bf16_tflops_sustained / 2 (TF32 rate is half the bf16 rate)'
                % peaks['source'],
    }

  total = B * world * args.steps
  value = total / (ms_total * 1e-3)
  e2e_value = total / (ms_e2e * 1e-3)
  line = {
      'metric': 'QM8 molecules/sec (forward)', 'value': value, 'unit': 'molecules/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'QM8 LanczosNet forward (config/qm8_lanczos_net.yaml), K=20, '
                             'batch=%d per GPU, N=26 padded, 7 layers, fp32 (3xTF32 tensor cores)' % B,
                 'global_batch': B * world, 'parallelism': 'dp%d' % world,
                 'cache': 'inputs larger than L2: %d distinct resident batches rotated '
                          '(%.0f MB > 126 MB L2), each copied into the CUDA graph\'s static input buffers' %
                          (NUM_BATCHES, NUM_BATCHES * h2d_bytes / 1e6)},
      'e2e': {'value': e2e_value, 'unit': 'molecules/s', 'h2d_bytes_per_step': h2d_bytes,
              'd2h_bytes_per_step': int(out_host.numel() * 4), 'ms_per_step': ms_e2e / args.steps},
      'gpu_launches': int(launches),
      'clocks': clocks,
      'roofline': roof,
      'oracle_check_max_abs_err': max_err,
  }
  if rank == 0:
    if not args.no_cpu_baseline and world == 1:
      sample = 256
      cb = make_batches(1, sample, 4242)[0]
      nthr = pick_cpu_threads(params, spec, cb)
      t = time_cpu_baseline(params, spec, cb, iters=3)
      line['cpu_baseline'] = {
          'value': sample / t, 'unit': 'molecules/s', 'cores': nthr, 'kind': 'port',
          'sample': '3 timed forwards of %d molecules (median), torch CPU fp32 oracle port of '
                    'model/lanczos_net.py, best of {4,8,16,32,all} threads = %d (host has %d cores)'
                    % (sample, nthr, os.cpu_count())}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
