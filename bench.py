#!/usr/bin/env python
"""Benchmark of the PdsNetwork.forward hot path (stereo pairs / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision fp16x2|fp32|bf16x3|bf16x2|bf16|fp16] [--workload C2|C3|C4|C1]
                    [--images u8|f32] [--streams 4] [--extra-configs C3,C4,C5]

One process per GPU (torchrun for N > 1: ranks are independent replicas, one
NCCL broadcast of the weights at start-up, NO collective in the timed region).
A step = one PdsNetwork.forward (eval) over one batch of synthetic stereo pairs.
Rank 0 prints ONE JSON line (see DESIGN.md "Measurement").

--impl reference times the reference's CPU implementation of the same path on
the host cores: the torch port under oracle/ (the reference itself is Python on
ATen operators and cannot travel to the GPU box; the port dispatches to the
same CPU kernels).  That leg, and cpu_baseline, are the only places this file
touches oracle/.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name: (H, W, maximum_disparity, description)
    'C1': (64, 128, 63, '128x64 md=63 (unit-test scale)'),
    'C2': (540, 960, 191, '960x540 D=192 (FlyingThings3D shape)'),
    'C3': (540, 960, 255, '960x540 D=256 (extended disparity range)'),
    'C4': (375, 1242, 191, '1242x375 D=192 (KITTI shape)'),
}
METRIC = 'stereo pairs/sec at 960x540 D=192'


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap,power.limit')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.samples.append([f.strip() for f in line.split(',')])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons, power, limit = [], 0, set(), [], None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                power.append(float(s[2]))
                limit = float(s[7]) if len(s) > 7 else limit
                for name, flag in zip(names, s[3:7]):
                    if flag.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm),
                # the step is energy-bound under the board's power cap: draw vs limit, watts
                'power_w': statistics.median(power) if power else None, 'power_limit_w': limit}


def synthetic_pairs(n, batch, H, W, device, seed=0, pinned=False, images='f32'):
    """`n` distinct seeded stereo pairs: right = left shifted by a few px + noise, 0..255.
    images: 'f32' = float (B,3,H,W) as the reference's data loader delivers them
    (dataset.py:67-72); 'u8' = the same images rounded, interleaved uint8 (B,H,W,3) as the decoder
    leaves them (the input path of INTEGRATION.md, a quarter of the upload)."""
    g = torch.Generator().manual_seed(seed)
    pairs = []
    for i in range(n):
        left = torch.rand(batch, 3, H, W, generator=g) * 255
        right = torch.rand(batch, 3, H, W, generator=g) * 255
        s = 4 + 3 * i
        right[..., :-s] = 0.8 * left[..., s:] + 0.2 * right[..., :-s]
        if images == 'u8':
            left = left.round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
            right = right.round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        if pinned:
            pairs.append((left.pin_memory(), right.pin_memory()))
        else:
            pairs.append((left.to(device), right.to(device)))
    return pairs


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
            'source': 'fallback'}


# 16-bit tensor-core products issued per reference (fp32) multiply-add by the split-operand
# precisions: the tensor pipe is busy `products` times the algorithmic FLOPs
PRODUCTS_PER_MAC = {'fp16x2': 3, 'bf16x2': 3, 'bf16x3': 6, 'bf16': 1, 'fp16': 1}


def kernel_rooflines(report, steps, peaks, precision='fp16x2'):
    """Per-kernel-class roofline from the live CUDA-event profile of `steps` steps.
    Algorithmic FLOPs / bytes per launch come from the library's own accounting
    (pds_profiler_read_work: reference FLOP count of the layer, minimum HBM bytes;
    DESIGN.md section 4).  `achieved`, `peak`, `frac` are always ALGORITHMIC work over the
    measured peak; a class is bound by the roof it is closer to, where the tensor roof counts
    the products actually issued (`executed_frac` = tensor_frac x products per reference MAC)."""
    out = []
    products = PRODUCTS_PER_MAC.get(precision, 1)
    for name, (launches, ms, flops, nbytes) in sorted(report.items(), key=lambda kv: -kv[1][1]):
        if launches == 0 or ms <= 0:
            continue
        sec = ms / 1e3
        entry = {'kernel': name, 'launches_per_step': launches / steps,
                 'ms_per_step': ms / steps, 'avg_us': ms / launches * 1e3}
        tensor = flops / sec / 1e12 / peaks['bf16_tflops_sustained'] if flops else 0.0
        hbm = nbytes / sec / 1e9 / peaks['hbm_gbs'] if nbytes else 0.0
        if flops or nbytes:
            entry['tflops'] = flops / sec / 1e12
            entry['gbs'] = nbytes / sec / 1e9
            entry['tensor_frac'], entry['hbm_frac'] = tensor, hbm
            on_tensor_cores = name.startswith('conv3x3_tc') or name.startswith('conv_tcg')
            executed = tensor * (products if on_tensor_cores else 1)
            if on_tensor_cores:
                entry['products_per_reference_mac'] = products
                entry['executed_frac'] = executed
            if executed >= hbm:
                entry.update(bound='tensor', unit='TFLOP/s', achieved=entry['tflops'],
                             peak=peaks['bf16_tflops_sustained'], frac=tensor)
            else:
                entry.update(bound='hbm', unit='GB/s', achieved=entry['gbs'],
                             peak=peaks['hbm_gbs'], frac=hbm)
        out.append(entry)
    return out


def cpu_forward_seconds(H, W, md, reps, threads=None):
    """Times the torch port of the reference forward on the host cores."""
    from oracle import synth, torch_port
    if threads:
        torch.set_num_threads(threads)
    params = {k: torch.from_numpy(v) for k, v in
              synth.make_params(synth.network_specs(), 61).items()}
    (left, right), = synthetic_pairs(1, 1, H, W, 'cpu')
    times = []
    with torch.no_grad():
        for _ in range(reps + 1):
            t0 = time.time()
            torch_port.network_forward(left, right, params, md)
            times.append(time.time() - t0)
    return times[1:] if reps else times      # first run is the warm-up


def run_reference(args, H, W, md, desc):
    rank = env_int('RANK', 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import synth, torch_port
    params = {k: torch.from_numpy(v) for k, v in
              synth.make_params(synth.network_specs(), 61).items()}
    # bounded sample: full pair unless the run would not end within a few minutes,
    # then a horizontal band of the pair (throughput scaled by the row ratio)
    rows, budget_s = H, 240.0
    (left, right), = synthetic_pairs(1, 1, H, W, 'cpu')
    with torch.no_grad():
        t0 = time.time()
        torch_port.network_forward(left, right, params, md)
        first = time.time() - t0
        total_steps = args.steps + args.warmup
        if first * total_steps > budget_s:
            rows = max(64, int(H * budget_s / (first * total_steps)) // 64 * 64)
            left, right = left[..., :rows, :].contiguous(), right[..., :rows, :].contiguous()
        for _ in range(args.warmup):
            torch_port.network_forward(left, right, params, md)
        t0 = time.time()
        for _ in range(args.steps):
            torch_port.network_forward(left, right, params, md)
        elapsed = time.time() - t0
    frac = rows / H
    value = args.steps * frac / elapsed
    sample = (f'{args.steps} x one full {W}x{H} pair' if rows == H else
              f'{args.steps} x a {W}x{rows} band of the pair (pairs/s scaled by {frac:.3f})')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': elapsed / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'batch_per_gpu': 1, 'maximum_disparity': md},
        'impl_config': {'implementation': 'torch port of the reference forward on host cores '
                                          '(oracle/torch_port.py, ATen CPU kernels)'},
        'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def time_pipeline(pipe, items, steps, barrier, max_over_ranks, out=None, download=True):
    """CUDA-event time (ms, max over ranks) of `steps` pairs through HostPipeline.run after an
    untimed pass of the same length (the caching allocator's per-stream pools reach their steady
    state: a first pass that keeps `steps` results alive pays cudaMalloc calls inside the region)."""
    fill = [items[i % len(items)] for i in range(steps)]
    warm = pipe.run(fill, out=out, download=download)
    barrier()
    del warm
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from practicaldeepstereo_nips2018_b200 import _capi, pipeline as pds_pipeline
    launches0 = _capi.launch_count() + pds_pipeline.replayed_launches()
    t0 = time.perf_counter()
    start.record()
    outs = pipe.run((items[i % len(items)] for i in range(steps)), out=out, download=download)
    stop.record()
    time_pipeline.launches = _capi.launch_count() + pds_pipeline.replayed_launches() - launches0   # kernels of the timed region (eager launches + kernels inside replayed graphs)
    time_pipeline.host_ms = (time.perf_counter() - t0) * 1e3       # host time to enqueue them
    barrier()
    del outs
    return max_over_ranks(start.elapsed_time(stop))


def sync_latency_ms(net, pairs, reps=12):
    """The reference's own protocol (trainer.py:141-148): wall time of ONE forward bracketed by
    torch.cuda.synchronize(); median of `reps` after the warm-up the caller has already done."""
    times = []
    for i in range(reps):
        left, right = pairs[i % len(pairs)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        net(left, right)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(times), min(times)


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the real stdout."""
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # Libraries talk on stdout (NCCL prints its version there when NCCL_DEBUG=VERSION is set in the
    # image): everything but the JSON line goes to stderr, so that stdout carries exactly one line.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('PDS_B200_PRECISION', 'fp16x2'))
    ap.add_argument('--workload', default='C2', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=1, help='stereo pairs per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--images', default='u8', choices=['f32', 'u8'],
                    help="host images of the e2e leg: interleaved uint8 (B,H,W,3) as the decoder leaves "
                         "them (default; dataset.py:67-72 converts on the host, here the kernel does) or "
                         "float (B,3,H,W) as the reference's loader yields; the other one is reported as "
                         "e2e_other_images")
    ap.add_argument('--preheat-s', type=float, default=4.0,
                    help='seconds of untimed forwards before the timed legs (sustained clocks for every leg)')
    ap.add_argument('--streams', type=int, default=4,
                    help='compute streams of the e2e serving pipeline (pairs dealt round-robin)')
    ap.add_argument('--graphs', type=int, default=1,
                    help='1: every compute stream replays a CUDA graph of the forward (pipeline.GraphedNetwork, '
                         'HostPipeline(graphs=True)); 0: eager kernel launches')
    ap.add_argument('--extra-configs', default='C3,C3-bf16,C4,C5',
                    help="further BASELINE.json configurations measured after the headline one and "
                         "reported under 'other_configs' of the same JSON line (C3-bf16 = C3 with plain "
                         "bf16 operands; C5 = 8 pairs per GPU per step at C2: batch 64 on 8 GPUs); '' disables")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    H, W, md, desc = WORKLOADS[args.workload]

    if args.impl == 'reference':
        run_reference(args, H, W, md, desc)
        return

    rank, world, local = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)
    assert torch.cuda.is_available(), 'bench.py --impl ours needs a CUDA device'
    from practicaldeepstereo_nips2018_b200 import parallel
    numa_cpus = parallel.bind_to_gpu_numa_node(local)   # before any pinned allocation (first touch)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from practicaldeepstereo_nips2018_b200 import PdsNetwork, _capi
    from practicaldeepstereo_nips2018_b200 import pipeline as pds_pipeline
    from practicaldeepstereo_nips2018_b200.pipeline import GraphedNetwork, HostPipeline

    torch.backends.cudnn.allow_tf32 = False        # embedding (cuDNN) stays fp32 like the oracle
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True          # as the reference trainer does (trainer.py:32-34)
    torch.manual_seed(0)
    net = PdsNetwork.default(md, precision=args.precision).to(dev).eval()
    if world > 1:
        parallel.broadcast_parameters(net, src=0)   # the only collective: weights, once

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    other_images = 'f32' if args.images == 'u8' else 'u8'
    pairs = synthetic_pairs(4, args.batch, H, W, dev, seed=1000 + rank)
    host_pairs = synthetic_pairs(4, args.batch, H, W, dev, seed=2000 + rank, pinned=True, images=args.images)
    host_pairs_other = synthetic_pairs(4, args.batch, H, W, dev, seed=2000 + rank, pinned=True, images=other_images)

    with torch.no_grad():
        for i in range(args.warmup):
            net(*pairs[i % len(pairs)])
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
            t_wait = time.time()
            while not sampler.samples and time.time() - t_wait < 3.0:   # nvidia-smi takes a moment to start
                net(*pairs[0])
                time.sleep(0.02)

        # ---- pre-heat: the value leg used to run on a GPU that had been idle (boost clocks, power
        # budget unspent) and the e2e leg on a warm one; at 8 ranks in one chassis that alone read as
        # "e2e = 0.965 x value" (tools/e2e_probe.py: the legs repeated in a loop agree within 1 %).
        # Both legs are now measured in the sustained state.
        t_heat = time.time()
        while time.time() - t_heat < args.preheat_s:
            for i in range(8):
                net(*pairs[i % len(pairs)])
            torch.cuda.synchronize()

        # ---- value: inputs resident in HBM ------------------------------------------
        # both numbers go through pipeline.HostPipeline, the package's serving call: pairs are
        # dealt round-robin to `--streams` compute streams (the latency-bound deep hourglass
        # layers of one pair overlap the other pairs' work); --streams 1 = plain back-to-back calls
        pipe = HostPipeline(net, dev, streams=args.streams, graphs=bool(args.graphs))
        ms = time_pipeline(pipe, pairs, args.steps, barrier, max_over_ranks, download=False)
        launches = time_pipeline.launches

        # ---- e2e: public API with HOST buffers, H2D + D2H inside the timed region ------
        # pipeline.HostPipeline is the package's serving call: per pair, upload of both images from
        # pinned memory (copy stream, overlapped with the previous pair's forward), forward,
        # download of the disparity map
        d2h = [torch.empty((args.batch, H, W), dtype=torch.float32).pin_memory() for _ in range(2 * max(1, args.streams))]
        e2e_ms = time_pipeline(pipe, host_pairs, args.steps, barrier, max_over_ranks, out=d2h)
        if sampler:
            sampler.stop()
        e2e_other_ms = time_pipeline(pipe, host_pairs_other, args.steps, barrier, max_over_ranks, out=d2h)

        # ---- back-to-back forwards on ONE stream, and the reference's sync-per-forward latency ----
        pipe1 = HostPipeline(net, dev, streams=1, graphs=bool(args.graphs))
        ms_1stream = time_pipeline(pipe1, pairs, args.steps, barrier, max_over_ranks, download=False)
        call = GraphedNetwork(net) if args.graphs else net
        for i in range(3):
            call(*pairs[i % len(pairs)])
        lat_median, lat_min = sync_latency_ms(call, pairs)
        lat_median = max_over_ranks(lat_median)

        # ---- per-kernel CUDA-event profile (separate instrumented pass) ----------------
        report = {}
        if rank == 0:
            _capi.profiler_reset()
            _capi.profiler_enable(True)
            for i in range(args.steps):
                net(*pairs[i % len(pairs)])
            torch.cuda.synchronize()
            _capi.profiler_enable(False)
            report = _capi.profiler_report()
        barrier()

        # ---- the other BASELINE.json configurations, same protocol, fewer steps -------------
        others = []
        del pairs, host_pairs, host_pairs_other, d2h

        def free_everything(*closables):
            import gc
            for q in closables:
                if hasattr(q, 'close'):
                    q.close()
            torch.cuda.synchronize()
            net.release_workspaces()
            gc.collect()
            torch.cuda.empty_cache()

        def run_other_config(name, odesc, oH, oW, omd, obatch, osteps, ostreams, net):
            net.set_maximum_disparity(omd)
            opairs = synthetic_pairs(2, obatch, oH, oW, dev, seed=3000 + rank)
            ohost = synthetic_pairs(2, obatch, oH, oW, dev, seed=4000 + rank, pinned=True, images=args.images)
            od2h = [torch.empty((obatch, oH, oW), dtype=torch.float32).pin_memory() for _ in range(2 * max(1, ostreams))]
            for i in range(3):
                net(*opairs[i % 2])
            opipe = HostPipeline(net, dev, streams=ostreams, graphs=bool(args.graphs))
            ocall = GraphedNetwork(net) if args.graphs else net
            try:
                oms = time_pipeline(opipe, opairs, osteps, barrier, max_over_ranks, download=False)
                oe2e = time_pipeline(opipe, ohost, osteps, barrier, max_over_ranks, out=od2h)
                for i in range(3):
                    ocall(*opairs[i % 2])
                olat, _ = sync_latency_ms(ocall, opairs, reps=10)
                olat = max_over_ranks(olat)
            finally:
                free_everything(opipe, ocall)
            total = osteps * obatch * world
            return {'config': name, 'workload': odesc, 'batch_per_gpu': obatch, 'maximum_disparity': omd,
                    'steps': osteps, 'streams_per_gpu': ostreams, 'value': total / (oms / 1e3), 'unit': 'pairs/s',
                    'ms_per_step': oms / osteps,
                    'e2e': {'value': total / (oe2e / 1e3), 'unit': 'pairs/s',
                            'h2d_bytes_per_step': 2 * obatch * 3 * oH * oW * (4 if args.images == 'f32' else 1),
                            'd2h_bytes_per_step': obatch * oH * oW * 4},
                    'latency_ms': olat}

        free_everything(pipe, pipe1, call)
        for name in [c for c in args.extra_configs.split(',') if c]:
            onet, oprecision = net, args.precision
            if name == 'C5':
                oH, oW, omd, odesc, obatch = 540, 960, 191, 'batch 64 of 960x540 D=192 over 8 GPUs = 8 pairs per GPU per step', 8
            elif name == 'C3-bf16':
                # BASELINE.json names this configuration with bf16 operands: single-term tensor-core
                # operands, NOT an fp32-grade mode (tests/test_gpu_network.py judges it on MAE / 3PE)
                oH, oW, omd, odesc = WORKLOADS['C3']
                odesc, obatch, oprecision = odesc + ', plain bf16 operands', 1, 'bf16'
                onet = PdsNetwork.default(omd, precision='bf16')
                onet.load_state_dict(net.state_dict())
                onet = onet.to(dev).eval()
            elif name in WORKLOADS and name != args.workload:
                oH, oW, omd, odesc = WORKLOADS[name]
                obatch = 1
            else:
                continue
            osteps = max(8, args.steps // (3 * obatch)) if obatch > 1 else max(12, args.steps // 3)
            ostreams = args.streams if obatch == 1 else 1     # eight pairs per step fill the GPU on one stream
            try:
                others.append(run_other_config(name, odesc, oH, oW, omd, obatch, osteps, ostreams, onet))
                others[-1]['precision'] = oprecision
            except Exception as exc:                           # a failed extra configuration never costs the headline line
                others.append({'config': name, 'workload': odesc, 'error': f'{type(exc).__name__}: {exc}'[:300]})
                free_everything()
            if onet is not net:
                onet.release_workspaces()
                del onet
                torch.cuda.empty_cache()
        net.set_maximum_disparity(md)
    barrier()

    if rank == 0:
        peaks = load_peaks()
        total_pairs = args.steps * args.batch * world
        kernels = kernel_rooflines(report, args.steps, peaks, args.precision)
        dominant = next((k for k in kernels if 'achieved' in k), None)
        roofline = None
        traffic = {}
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        if dominant:
            roofline = {'bound': dominant['bound'], 'achieved': dominant['achieved'],
                        'peak': dominant['peak'], 'unit': dominant['unit'],
                        'frac': dominant['frac'], 'tensor_frac': dominant['tensor_frac'],
                        'executed_frac': dominant.get('executed_frac'),
                        'products_per_reference_mac': dominant.get('products_per_reference_mac'),
                        'note': 'achieved/frac count the reference (fp32) FLOPs of the layer; the fp32-grade '
                                'split-operand precision issues products_per_reference_mac 16-bit tensor-core '
                                'products per reference multiply-add, executed_frac is the share of the '
                                'measured bf16 peak those occupy',

                        'hbm_frac': dominant['hbm_frac'], 'tflops': dominant['tflops'], 'gbs': dominant['gbs'],
                        'traffic': traffic.get(dominant['kernel']),
                        'traffic_source': traffic.get('_source') if dominant['kernel'] in traffic else None,
                        'kernel': dominant['kernel'],
                        'peak_source': peaks['source'],
                        'timing': 'CUDA events around every launch, separate instrumented pass of '
                                  f'{args.steps} steps'}
        image_bytes = {'f32': 4, 'u8': 1}
        line = {
            'metric': METRIC, 'value': total_pairs / (ms / 1e3), 'unit': 'pairs/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': (total_pairs / (ms / 1e3)) / (1.0 / 0.62) if args.workload == 'C2' else None,
            'dtype': {'fp32': 'f32', 'bf16': 'bf16'}.get(args.precision, args.precision),
            'data': 'synthetic',
            # `config` names the workload only (identical in the reference arm); how this arm runs it
            # is under `impl_config`
            'config': {'workload': desc, 'batch_per_gpu': args.batch, 'maximum_disparity': md},
            'impl_config': {'precision': args.precision, 'parallelism': f'replicas x{world}',
                            'streams_per_gpu': args.streams, 'host_images': args.images,
                            'cuda_graphs': bool(args.graphs), 'preheat_s': args.preheat_s,
                            'numa_bound_cpus': len(numa_cpus) if numa_cpus else None,
                            'l2': 'per-step working set > 1 GB (>> 126 MB L2); 4 rotating input pairs',
                            'embedding': ('own tcgen05 kernels' if args.precision != 'fp32'
                                          else 'ATen/cuDNN fp32 (TF32 off)')},
            'e2e': {'value': total_pairs / (e2e_ms / 1e3), 'unit': 'pairs/s',
                    'h2d_bytes_per_step': 2 * args.batch * 3 * H * W * image_bytes[args.images],
                    'd2h_bytes_per_step': args.batch * H * W * 4},
            'e2e_other_images': {'host_images': other_images, 'value': total_pairs / (e2e_other_ms / 1e3),
                                 'unit': 'pairs/s',
                                 'h2d_bytes_per_step': 2 * args.batch * 3 * H * W * image_bytes[other_images],
                                 'd2h_bytes_per_step': args.batch * H * W * 4},
            # back-to-back forwards on one stream (no overlap between pairs), and the reference's own
            # protocol: one forward bracketed by cuda.synchronize() (trainer.py:141-148), median of 12
            'value_1stream': total_pairs / (ms_1stream / 1e3),
            'latency_ms': lat_median, 'latency_ms_min': lat_min,
            'gpu_launches': launches,
            'clocks': sampler.summary() if sampler else None,
            'roofline': roofline,
            'other_configs': others,
            'kernels': kernels,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            times = cpu_forward_seconds(H, W, md, reps=10, threads=cores)    # ~10 s of CPU work
            line['cpu_baseline'] = {'value': 1.0 / statistics.median(times), 'unit': 'pairs/s',
                                    'cores': cores, 'kind': 'port',
                                    'sample': f'{len(times)} x one full {W}x{H} pair, torch port of '
                                              'the reference forward (after 1 warm-up)'}
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
