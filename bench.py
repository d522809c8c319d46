#!/usr/bin/env python
"""Benchmark of the NeuBE generator-forward hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 256]

One "step" = one pass of the hot path over one batch of B = 256 synthetic 128x128 patches with a distinct z per
patch (BASELINE.json configs[1]): geometry encoder -> mapping -> synthesis (grouped modulated convs on tcgen05)
-> ToRGB/triad -> triband composite -> uint8 tiles.  `value` times it with inputs resident in HBM; `e2e` times
the same work through `TriadPaintEngine.render_patches_host` with pinned HOST buffers (H2D + D2H inside the
timed region).  Multi-GPU (torchrun, one rank per GPU): every rank renders its own B patches (independent units,
no data-path collective) -> weak scaling; time = max over ranks.

`--impl reference` times the CPU arm: the oracle port of the reference's own CPU path (`oracle/neube_oracle.py`,
validated against the unmodified reference in the build container), all host threads, on a bounded sample of the
same workload.  /root/reference does not exist on the GPU box, so the port is the only reference that can run there.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from brushstroke_engine_b200 import params as P          # noqa: E402
from brushstroke_engine_b200 import synthetic             # noqa: E402

METRIC = 'stroke_patches_per_sec_128x128'
UNIT = 'patches/s'


def workload(batch: int, n_sets: int = 8):
    """BASELINE config 2: z_i = RandomState(i).randn(1,64); geometry = crops at stride 88 from the synthetic
    2000^2 drawing; positions = crop (y, x).  `n_sets` distinct batches are rotated so that
    consecutive steps never re-read the same input from L2."""
    from brushstroke_engine_b200.stylizer import generate_stitching_crops, pad_geo
    guidance = synthetic.synthetic_guidance(2000, 2000, num_lines=64, seed=0)
    geom = pad_geo(guidance, 10)
    crops, geom = generate_stitching_crops(geom, 128, 'all', 20)
    sets = []
    for s in range(n_sets):
        idx = [(s * batch + i) % len(crops) for i in range(batch)]
        patches = np.stack([geom[crops[i][0]:crops[i][0] + 128, crops[i][1]:crops[i][1] + 128, 0] for i in idx])
        pos = np.array([(crops[i][0], crops[i][1]) for i in idx], dtype=np.int64)
        z = np.concatenate([np.random.RandomState(seed=s * batch + i).randn(1, 64) for i in range(batch)])
        sets.append((torch.from_numpy(patches), torch.from_numpy(z), torch.from_numpy(pos)))
    return sets


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons of one GPU, sampled while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                'reasons': sorted(reasons), 'samples': len(self.samples)}


def cpu_port_patches_per_sec(sets, n_patches: int, reps: int, threads: int):
    """The oracle port (fp32, = the reference's force_fp32 CPU path) on the first `n_patches` patches of the workload."""
    from oracle import neube_oracle as O
    torch.set_num_threads(threads)
    cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
    gp = P.init_generator_params(cfg, 0, 0.1)
    ep = P.init_encoder_params(ecfg, 1, 0.1)
    patches, z, pos = sets[0]
    geom = 1 - (255 - patches[:n_patches].to(torch.float32)) / 255.0
    geom = geom[:, None]
    times = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            gf = O.geometry_encode(ep, ecfg, geom)
            _, dbg = O.generator_forward(gp, cfg, z[:n_patches], gf, positions=pos[:n_patches])
            rgba = O.triad_composite(dbg['uvs'], dbg['colors'], 'clear')
            O.to_uint8_tile(rgba, 10)
            times.append(time.perf_counter() - t0)
    return n_patches / float(np.median(times)), times


def cpu_port_b1_ms(sets, threads: int) -> float:
    """BASELINE configs[0]: the CPU path on ONE 128x128 patch, batch 1: median of 20 runs after 3 warm-ups (ms)."""
    times = []
    for _ in range(3):
        cpu_port_patches_per_sec(sets, 1, 1, threads)
    _, times = cpu_port_patches_per_sec(sets, 1, 20, threads)
    return float(np.median(times)) * 1e3


def gpu_incumbent(sets_dev, dev, B, steps=4):
    """The incumbent GPU path on the same batch (tools/incumbent_torch.py: the reference as shipped on this torch = its
    impl='ref' ops + cuDNN): patches/s for stock mixed fp16 (channels_last), the same with bf16
    operands, and forced fp32.  CUDA events, 2 warm-ups + `steps` timed steps per mode."""
    sys.path.insert(0, os.path.join(REPO, 'tools'))
    from incumbent_torch import Incumbent
    cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
    gp = P.init_generator_params(cfg, 0, 0.1)
    ep = P.init_encoder_params(ecfg, 1, 0.1)
    out = {}
    for tag, lowp, f32 in (('fp16_channels_last', torch.float16, False), ('bf16_channels_last', torch.bfloat16, False), ('fp32', torch.float16, True)):
        try:
            inc = Incumbent(gp, ep, cfg, ecfg, dev, lowp=lowp, force_fp32=f32)
            with torch.no_grad():
                for i in range(2):
                    g, z, pos = sets_dev[i % len(sets_dev)]
                    inc.render_tiles(g, z, pos)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    g, z, pos = sets_dev[i % len(sets_dev)]
                    inc.render_tiles(g, z, pos)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = {'patches_per_s': B / ms * 1e3, 'ms_per_step': ms}
            del inc
            torch.cuda.empty_cache()
        except Exception as ex:                                      # an out-of-memory cuDNN plan must not take the bench line down
            out[tag] = {'error': f'{type(ex).__name__}: {str(ex)[:160]}'}
            torch.cuda.empty_cache()
    out['what'] = ('reference as shipped on torch 2.x (plugins do not load -> impl=ref ATen ops + cuDNN convs), restated in '
                   'tools/incumbent_torch.py; same batch, encoder + generator + composite, device-resident inputs')
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sets = workload(args.batch, n_sets=1)
    n = 8
    for _ in range(args.warmup):
        cpu_port_patches_per_sec(sets, n, 1, threads)
    v, times = cpu_port_patches_per_sec(sets, n, max(args.steps, 1), threads)
    b1_ms = cpu_port_b1_ms(sets, threads)
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': float(np.median(times)) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'NeuBE style2 generator forward (encoder+mapping+synthesis+triad composite), {n}-patch sample per step '
                                   f'(the first {n} patches of the batch-{args.batch} workload of BASELINE configs[1]; the CPU path is '
                                   f'fastest per patch at this batch size), 128x128 patches, distinct z per patch',
                       'batch': n, 'sample_of_batch': args.batch},
            'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': f'{n} patches per step (first {n} of the {args.batch}-patch batch), fp32, {threads} threads',
                             'config0_b1_ms': b1_ms,
                             'config0': 'BASELINE configs[0]: one 128x128 patch, batch 1, median of 20 after 3 warm-ups'},
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-canvas', action='store_true')
    ap.add_argument('--no-blend', action='store_true', help='skip the feature-blending (level 2) canvas leg')
    ap.add_argument('--canvas', type=int, default=4096, help='side of the synthetic canvas for the stylization leg')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer leg (profiling runs under ncu)')
    ap.add_argument('--no-incumbent', action='store_true', help='skip the cuDNN / ATen incumbent leg')
    ap.add_argument('--eager', action='store_true', help='issue every kernel of the step from Python instead of replaying its CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    from brushstroke_engine_b200 import _lib
    from brushstroke_engine_b200.engine import GanBrushOptions, TriadPaintEngine
    cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
    gp = P.init_generator_params(cfg, 0, 0.1)
    ep = P.init_encoder_params(ecfg, 1, 0.1)
    engine = TriadPaintEngine(gp, ep, dev, mode='bf16')
    B = args.batch
    sets = workload(B)
    # device-resident copies (for `value`) and pinned host copies (for `e2e`)
    dsets = []
    for patches, z, pos in sets:
        g = (1 - (255 - patches.to(dev).to(torch.float32)) / 255.0)[:, None].contiguous()
        dsets.append((g, z.to(dev), pos.to(dev)))
    hsets = [(p_.pin_memory(), z.pin_memory(), pos.pin_memory()) for p_, z, pos in sets]
    out_host = torch.empty((B, 108, 108, 4), dtype=torch.uint8, pin_memory=True)

    # the step runs as ONE CUDA-graph replay (engine.BatchSession) + the last synthesis layer and the composite launched
    # eagerly after it, so that the dominant kernel can be bracketed by CUDA events inside the timed region; --eager issues
    # every launch from Python instead (A/B)
    bsess = None
    if not args.eager:
        from brushstroke_engine_b200.engine import BatchSession
        with torch.no_grad():
            bsess = BatchSession(engine, B, 10, split_last_layer=True)

    def step(i):
        g, z, pos = dsets[i % len(dsets)]
        if bsess is not None:
            return bsess.run(g, z, pos)
        opts = GanBrushOptions()
        opts.set_style(z)
        opts.position = pos
        tiles, _ = engine.render_tiles(g, opts, crop_margin=10)
        return tiles

    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]

    def e2e_step(i):
        # two pinned result buffers: the download of batch i overlaps the compute of batch i+1; every result is complete on
        # the host (event.synchronize) before its buffer is handed out again
        p_, z, pos = hsets[i % len(hsets)]
        return engine.render_patches_host(p_, z, pos, crop_margin=10, out=out_hosts[i & 1], wait=False)[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            step(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        DOM = 'b128.conv1'
        engine.G.probe = {DOM: []}
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        launches = _lib.launch_count() - launches0
        if bsess is not None:
            launches += bsess.kernels_per_replay * args.steps          # kernels inside the replayed graph + the eager tail counted above
        ms_total = e0.elapsed_time(e1)
        probe = engine.G.probe[DOM]
        engine.G.probe = None
        dom_ms = float(np.mean([a.elapsed_time(b) for a, b in probe]))
        # the clock sampler forks nvidia-smi every 100 ms, which competes with the launching thread for host cores and the
        # driver lock: it covers the device-timed region above (what the clocks line is about) and stops before the
        # host-driven legs
        sampler.stop_flag.set()
        sampler.join(timeout=2)
        # ---- end-to-end (host buffers): median of 3 repetitions of `steps` steps (wall clock, host-driven) ----
        e2e_s = float('nan')
        if not args.no_e2e:
            for i in range(3):
                e2e_step(i).synchronize()
            reps = []
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                pending = None
                for i in range(args.steps):
                    ev = e2e_step(i)
                    if pending is not None:
                        pending.synchronize()                           # batch i-1 is on the host
                    pending = ev
                pending.synchronize()
                barrier()
                reps.append(time.perf_counter() - t0)
            e2e_s = sorted(reps)[1]

    incumbent = None
    if world == 1 and not args.no_incumbent:
        incumbent = gpu_incumbent(dsets, dev, B)

    # ---- BASELINE configs[4] / configs[3]: 4096^2 canvas with the style interpolated across 8 z anchors, and the 2000^2
    #      line-drawing stylization with one style; crop rows sharded over the ranks, ONE batched send/recv of the owned canvas
    #      bands to rank 0 (time = guidance in -> finished uint8 canvas on rank 0) ----
    canvas_ms = None
    canvas_legs = {}
    if not args.no_canvas:
        from brushstroke_engine_b200 import stylizer

        def canvas_leg(size, interpolate, blended_too=False):
            guidance = synthetic.synthetic_guidance(size, size, num_lines=256 if size >= 4096 else 64, seed=0)
            job_crops, _ = stylizer.generate_stitching_crops(stylizer.pad_geo(guidance, 10), 128, 'all', 20)
            copts = GanBrushOptions()
            z_pp = None
            if interpolate:
                anchors = np.concatenate([np.random.RandomState(seed=k).randn(1, 64) for k in range(8)])
                xs = np.array([c[1] for c in job_crops], dtype=np.float64) / max(1, max(c[1] for c in job_crops))
                t_ = xs * 7.0
                k0 = np.clip(np.floor(t_).astype(int), 0, 6)
                a_ = (t_ - k0)[:, None]
                z_pp = torch.from_numpy((1 - a_) * anchors[k0] + a_ * anchors[k0 + 1]).to(dev)   # z = alpha z1 + (1 - alpha) z2 per patch
                copts.set_style(z_pp[:1])
            else:
                copts.set_style(torch.from_numpy(np.random.RandomState(594).randn(1, 64)).to(dev), '594')
            ctimes, htimes = [], []
            d_guidance = torch.from_numpy(guidance).to(dev)
            out = None
            with torch.no_grad():
                for rep in range(4):
                    barrier()
                    t0 = time.perf_counter()
                    out = stylizer.stylize(engine, d_guidance, copts, crop_margin=10, batch_size=B, z_per_patch=z_pp, to_host=False)
                    barrier()
                    if rep > 0:
                        ctimes.append((time.perf_counter() - t0) * 1e3)
                for rep in range(3):
                    barrier()
                    t0 = time.perf_counter()
                    stylizer.stylize(engine, guidance, copts, crop_margin=10, batch_size=B, z_per_patch=z_pp)
                    barrier()
                    if rep > 0:
                        htimes.append((time.perf_counter() - t0) * 1e3)
                # per-phase breakdown (its own run: filling `timings` synchronises between phases), max over ranks below
                phases = {}
                barrier()
                stylizer.stylize(engine, d_guidance, copts, crop_margin=10, batch_size=B, z_per_patch=z_pp, to_host=False, timings=phases)
                barrier()
                # the sharded canvas must equal the canvas one GPU renders by itself, bit for bit
                equal = None
                if world > 1 and rank == 0:
                    solo = stylizer.stylize(engine, d_guidance, copts, crop_margin=10, batch_size=B, z_per_patch=z_pp, to_host=False,
                                            distributed=False)
                    equal = bool(torch.equal(solo, out))
                barrier()
            names = ('setup', 'render', 'place', 'exchange', 'finish')
            pt = torch.tensor([phases.get(k, 0.0) for k in names] + [float(np.median(ctimes)), float(np.median(htimes))],
                              dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(pt, op=dist.ReduceOp.MAX)
            leg = {'size': size, 'patches': len(job_crops), 'ms': float(pt[5]), 'ms_host_to_host': float(pt[6]), 'n_gpus': world,
                   'phases_ms_max_over_ranks': {k: round(float(pt[i]), 3) for i, k in enumerate(names)},
                   'style': '8-anchor z interpolation along x' if interpolate else 'one style (seed 594)'}
            if equal is not None:
                leg['canvas_equals_1gpu'] = equal
            if blended_too and not args.no_blend:
                # the same drawing the way scripts/neube_stylize.sh renders it: --feature_blending_level=2 (one canvas over all ranks)
                btimes_, bout = [], None
                with torch.no_grad():
                    for rep in range(3):
                        barrier()
                        t0 = time.perf_counter()
                        bout = stylizer.stylize(engine, d_guidance, copts, crop_margin=10, batch_size=B, z_per_patch=z_pp, to_host=False,
                                                feature_blending_level=2)
                        barrier()
                        if rep > 0:
                            btimes_.append((time.perf_counter() - t0) * 1e3)
                    beq = None
                    if world > 1 and rank == 0:
                        solo = stylizer.stylize(engine, d_guidance, copts, crop_margin=10, batch_size=B, z_per_patch=z_pp, to_host=False,
                                                feature_blending_level=2, distributed=False)
                        beq = bool(torch.equal(solo, bout))
                    barrier()
                bt_ = torch.tensor([float(np.median(btimes_))], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(bt_, op=dist.ReduceOp.MAX)
                leg['ms_feature_blending_level2'] = float(bt_[0])
                if beq is not None:
                    leg['blended_equals_1gpu'] = beq
            return leg, d_guidance, copts, z_pp

        size = args.canvas
        leg, d_guidance, copts, z_pp = canvas_leg(size, True)
        canvas_legs['main'] = leg
        canvas_ms = leg['ms']
        canvas_host_ms = leg['ms_host_to_host']
        n_canvas_patches = leg['patches']
        # the reference's own stylization script runs with --feature_blending_level=2 (scripts/neube_stylize.sh): patches then
        # depend on their raster predecessors through the blend at 64^2 (stylizer._stylize_blended_phased)
        blend_ms = None
        blend_sharded_ms, blend_equal = None, None
        if not args.no_blend:
            # first with N ranks rendering N canvases (replicas, no collective: canvases per second), then ONE canvas over all ranks
            btimes = []
            with torch.no_grad():
                for rep in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    stylizer.stylize(engine, d_guidance, copts, crop_margin=10, feature_blending_level=2, z_per_patch=z_pp, to_host=False,
                                     distributed=False)
                    torch.cuda.synchronize()
                    if rep > 0:
                        btimes.append((time.perf_counter() - t0) * 1e3)
            bt = torch.tensor([float(np.median(btimes))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            blend_ms = float(bt[0])
            # ONE blended canvas on all ranks: the layers before / after the blend point sharded over the ranks, the blend on rank 0
            # (stylizer._stylize_blended_phased); rank 0 compares the bytes with its own single-GPU canvas
            if world > 1:
                stimes = []
                with torch.no_grad():
                    for rep in range(3):
                        barrier()
                        t0 = time.perf_counter()
                        shard_canvas = stylizer.stylize(engine, d_guidance, copts, crop_margin=10, feature_blending_level=2, z_per_patch=z_pp,
                                                        to_host=False)
                        barrier()
                        if rep > 0:
                            stimes.append((time.perf_counter() - t0) * 1e3)
                    if rank == 0:
                        solo = stylizer.stylize(engine, d_guidance, copts, crop_margin=10, feature_blending_level=2, z_per_patch=z_pp,
                                                to_host=False, distributed=False)
                        blend_equal = bool(torch.equal(shard_canvas, solo))
                        del solo
                st_ = torch.tensor([float(np.median(stimes))], dtype=torch.float64, device=dev)
                dist.all_reduce(st_, op=dist.ReduceOp.MAX)
                blend_sharded_ms = float(st_[0])
        barrier()
        del d_guidance
        if size != 2000:
            canvas_legs['config3_2000'] = canvas_leg(2000, False, blended_too=True)[0]          # BASELINE configs[3]

    # ---- interactive use: one 128^2 stroke patch per call through the reference-facing render_stroke
    #      (host uint8 patch in, host uint8 RGBA out; wall clock, rank 0) ----
    interactive_ms = None
    interactive_graph_ms = None
    interactive_batched_ms = None
    if rank == 0 and not args.no_e2e:
        from brushstroke_engine_b200.engine import GanBrushOptions as _GBO
        patch = np.ascontiguousarray(((1.0 - synthetic.synthetic_patch(128, seed=5)[0, 0]) * 255).astype(np.uint8)[:, :, None])   # [W,W,1], 255 = stroke
        iopts = _GBO()
        iopts.set_style(torch.from_numpy(np.random.RandomState(594).randn(1, 64)).to(dev), '594')
        iopts.position = torch.tensor([[88, 176]], device=dev)
        lat = []
        with torch.no_grad():
            for i in range(60):
                t0 = time.perf_counter()
                engine.render_stroke(patch, None, iopts)
                if i >= 10:
                    lat.append((time.perf_counter() - t0) * 1e3)
        interactive_ms = float(np.median(lat))
        # the same through a CUDA-graph session (one capture per brush, one replay per stroke patch)
        sess = engine.interactive_session(iopts, crop_margin=0)
        glat = []
        for i in range(110):
            t0 = time.perf_counter()
            sess.render_stroke(patch, (88 + i, 176))
            if i >= 10:
                glat.append((time.perf_counter() - t0) * 1e3)
        interactive_graph_ms = float(np.median(glat))
        # 32 concurrent sessions, one stroke each: one batched forward (server.StrokeBatcher) vs 32 single-patch calls
        from brushstroke_engine_b200 import server
        batcher = server.StrokeBatcher(engine)
        sessions = [server.DrawingSession(engine, style_seed=k, batcher=batcher) for k in range(32)]
        for s_ in sessions:
            s_.on_message(json.dumps({'type': 'set_option', 'option': 'positions', 'value': True}))
        rgba = np.zeros((128, 128, 4), dtype=np.uint8)
        rgba[..., 3] = patch[..., 0]
        blat = []
        with torch.no_grad():
            for i in range(25):
                reqs = [server.encode_render_request(rgba, 64 + k, 32 + i, 10) for k in range(32)]
                t0 = time.perf_counter()
                for s_, m_ in zip(sessions, reqs):
                    s_.on_message(m_)
                answers = server.DrawingSession.flush_all(batcher, sessions)
                if i >= 5:
                    blat.append((time.perf_counter() - t0) * 1e3)
                assert all(len(answers[s_]) == 1 for s_ in sessions)
        interactive_batched_ms = float(np.median(blat))
    barrier()
    times = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(times[0]), float(times[1])
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    # the dominant kernel is timed inside a step that lasts a few ms at full clock, not inside a seconds-long power-capped
    # run: the BURST peak is the denominator of `frac`; the sustained one is printed beside it
    peak_tf = peaks.get('bf16_tflops', 1633.0)
    peak_sus = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'MEASURED_PEAKS.json bf16_tflops (burst, measured)' if peaks else 'fallback 1.63 PFLOP/s burst (B200_PROFILING.md)'
    R, C = 128, 128
    dom_flops = 2.0 * C * C * 9 * R * R * B                          # algorithmic: 2*Cout*Cin*k^2*H*W per patch
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    step_flops = (8.676e9 + 1.73e9) * B                              # generator 8.676 + encoder 1.73 GFLOP per patch
    step_tf = step_flops / (ms_per_step * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(REPO, 'profiles', 'dominant_kernel_traffic.json')
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get('dram_bytes_per_launch')
        except Exception:
            traffic = None

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'NeuBE style2 generator forward (encoder+mapping+synthesis+triad composite), batch {B} '
                                   f'x 128x128 patches, distinct z per patch (BASELINE configs[1])',
                       'batch_per_gpu': B, 'parallelism': f'patch-sharded x{world} (no data-path collective)',
                       'l2': 'per-step working set ~6 GB >> 126 MB L2; inputs rotate over 8 distinct batches (8 x 16.8 MB)',
                       'mode': 'bf16 tensor-core (tcgen05), fp32 accumulate',
                       'launch': 'eager launches from Python' if bsess is None else
                                 f'one CUDA-graph replay per step ({bsess.kernels_per_replay} kernels) + last layer and composite eager (events around the dominant kernel)'},
            'roofline': {'bound': 'tensor', 'kernel': f'conv_tc_row128_kernel @ {DOM} (3x3 modconv 128->128 @128^2 with ToRGB/triad fused in the epilogue, batch {B})',
                         'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
                         'traffic': traffic, 'peak_source': peak_src, 'avg_launch_ms': dom_ms,
                         'algorithmic_flops_per_launch': dom_flops,
                         'peak_sustained': peak_sus, 'frac_of_sustained': achieved / peak_sus,
                         'step': {'what': 'whole step: algorithmic FLOPs of encoder + generator (10.406 GFLOP/patch) / ms_per_step',
                                  'achieved': step_tf, 'frac': step_tf / peak_tf, 'frac_of_sustained': step_tf / peak_sus}},
            'e2e': {'value': e2e_value, 'unit': UNIT,
                    'h2d_bytes_per_step': int(B * 128 * 128 + B * 64 * 8 + B * 2 * 8), 'd2h_bytes_per_step': int(B * 108 * 108 * 4),
                    'timing': f'median of 3 repetitions of {args.steps} steps, wall clock around render_patches_host with pinned host buffers'},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
        }
        if interactive_ms is not None:
            line['interactive'] = {'ms_per_stroke_patch': interactive_ms, 'ms_per_stroke_patch_cuda_graph': interactive_graph_ms,
                                   'ms_per_32_sessions_batched': interactive_batched_ms,
                                   'batched': '32 sessions x one binary render request each (wire decode -> ONE batched forward -> 32 encoded '
                                              'responses), wall clock per round (server.StrokeBatcher)',
                                   'what': 'batch 1, host uint8 patch -> host uint8 RGBA, wall-clock median (rank 0): TriadPaintEngine.render_stroke (eager, ~60 launches) and InteractiveSession.render_stroke (one CUDA-graph replay)'}
        if canvas_ms is not None:
            line['canvas'] = dict(canvas_legs['main'])
            line['canvas'].update({'ms_feature_blending_level2_1gpu': blend_ms,
                                   'blended_canvases_per_s': (world / (blend_ms * 1e-3)) if blend_ms else None,
                                   'ms_feature_blending_level2_sharded': blend_sharded_ms,
                                   'blended_sharded_equals_1gpu': blend_equal,
                                   'blended': 'feature_blending_level=2 (what scripts/neube_stylize.sh runs): every patch\'s layers before / '
                                              'after the blend point at batch 256, only the blend kernel in wavefront order '
                                              '(stylizer._stylize_blended_phased). _1gpu: one canvas per GPU (max over ranks; canvases/s = '
                                              'n_gpus / that); _sharded (n_gpus > 1): ONE canvas, those layers sharded over the ranks, the '
                                              'blend on rank 0, wall clock between barriers',
                                   'what': 'uint8 guidance on the device -> crops -> encoder+generator+composite (8-anchor z interpolation) -> '
                                           'every rank places its tiles into the canvas rows it owns -> one batched NCCL send/recv of those bands '
                                           'into the canvas on rank 0 (SURVEY 8d config 5); ms_host_to_host: host guidance in (each rank uploads '
                                           'only the rows its crops read), pinned host canvas out (wall clock, max over ranks)'})
            if 'config3_2000' in canvas_legs:
                line['canvas_2000'] = canvas_legs['config3_2000']
        if world == 1 and not args.no_incumbent:
            line['gpu_incumbent'] = incumbent
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            # bounded sample, ~10 s of CPU work: 8-patch batches (the port's fastest batch size; larger ones are slower per patch)
            n, reps = 8, 48
            v, times_cpu = cpu_port_patches_per_sec(sets, n, reps, threads)
            b1_ms = cpu_port_b1_ms(sets, threads)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'config0_b1_ms': b1_ms,
                                    'config0': 'BASELINE configs[0]: one 128x128 patch, batch 1, median of 20 after 3 warm-ups',
                                    'sample': f'{n} patches x {reps} reps of the same workload (fp32 oracle port, {threads} threads, median; '
                                              f'{sum(times_cpu):.1f} s of CPU work)'}
        print(json.dumps(line), flush=True)                             # stdout is a block-buffered file under the driver: flush before any teardown
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:                                               # noqa: BLE001 -- the line is out; a teardown hiccup must not cost the run
            pass


if __name__ == '__main__':
    main()
