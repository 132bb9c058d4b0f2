#!/usr/bin/env python
"""Benchmark of the NeRFool / IBRNet per-ray hot path (BASELINE.json: "rays/s fwd and PGD attack iters/s,
378x504 view ...").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] - the view-specific PGD attack on a synthetic
378x504 LLFF-shaped scene, 4 source views, 64 coarse + 64 importance samples.  One *step* is one attack
iteration of the hot path over ALL 190,512 rays of the target view: render_rays (coarse + fine) -> masked
MSE -> backward to the two source feature maps (the cuDNN encoder that turns that gradient into
delta.grad is out of scope and not timed here) -> sign-step on the feature maps (stand-in for the delta
update, so consecutive steps depend on each other).  With N > 1 each rank renders its own target view
(weak scaling, BASELINE configs[2] layout) and ONE NCCL allreduce sums the feature-map gradients.

value       rays/s through forward + backward, inputs resident in HBM, CUDA-event timed, max over ranks
e2e         the same step through the public API with the step's rays / target colours copied from pinned
            host memory and the loss read back, inside the timed region
roofline    the dominant kernel (largest share of the step) against the measured HBM peak, algorithmic
            gather/scatter bytes (560 B per (sample, view) row gathered, 512 B scattered);
            "kernels" carries the same for all four IBRNet kernels, with the dense-FLOP rate against the
            measured bf16 tensor peak (roofline_tensor)
cpu_baseline / --impl reference : the CPU oracle (oracle/ibrnet_oracle.py, a port of the reference path)
            on a bounded ray sample with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

H, W = 378, 504
SCENE_KIND = 'llff'
N_SAMPLES, N_IMPORTANCE = 64, 64
GATHER_B, SCATTER_B = 560, 512          # algorithmic bytes per (sample, view) row, SURVEY.md 8(d)


def load_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def load_tensor_peak():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f).get('bf16_tflops_sustained', 0) or 0) or None
    return 1381.5      # fallback: sustained dense bf16 of this pool (B200_PROFILING.md)


def load_traffic_table():
    """DRAM bytes per unit measured ONCE with `ncu --set full` (profiles/traffic_r01.json, source named inside)."""
    path = os.path.join(REPO, 'profiles', 'traffic_r01.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(',') for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for n, val in zip(names, r[5:9]):
                    if val.strip().lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {'sm_mhz': statistics.median(busy) if busy else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------
def build_workload(device, rank, V, seed=0):
    """Scene + nets, identical on every rank except the target view (rank r renders target r)."""
    from nerfool_b200.synthetic import make_scene, rays_for_view
    from nerfool_b200.mlp_network import IBRNet
    from nerfool_b200.projection import Projector
    world = int(os.environ.get('WORLD_SIZE', '1'))
    scene = make_scene(H, W, V, seed=seed, kind=SCENE_KIND, n_targets=max(world, 1))
    ray_o, ray_d = rays_for_view(scene['camera'][rank], H, W)
    args = types.SimpleNamespace(anti_alias_pooling=1)
    torch.manual_seed(seed)
    net_c = IBRNet(args, 32, N_SAMPLES)
    net_f = IBRNet(args, 32, N_SAMPLES + N_IMPORTANCE)
    with torch.no_grad():
        for n in (net_c, net_f):
            n.out_geometry_fc[2].bias += 0.3       # non-trivial compositing weights
    model = types.SimpleNamespace(net_coarse=net_c.to(device).eval(), net_fine=net_f.to(device).eval())
    host = {'ray_o': ray_o.pin_memory(), 'ray_d': ray_d.pin_memory(), 'rgb': scene['rgb'][rank].contiguous().pin_memory()}
    static = {'depth_range': scene['depth_range'].to(device), 'camera': scene['camera'][rank:rank + 1].to(device),
              'src_rgbs': scene['src_rgbs'].to(device), 'src_cameras': scene['src_cameras'].to(device)}
    featmaps = [f.to(device).contiguous() for f in scene['featmaps']]
    return scene, model, Projector(device), host, static, featmaps


def run_ours(a):
    import torch.distributed as dist
    from nerfool_b200 import _lib
    from nerfool_b200.attack import pgd_hot_step
    from nerfool_b200.render_ray import render_rays
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: nerfool_b200 has no CPU path (use --impl reference for the CPU baseline)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    group = None
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'      # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group('nccl', device_id=device)
        group = dist.group.WORLD
    _lib.load()
    scene, model, projector, host, static, featmaps = build_workload(device, rank, a.views)
    R = host['ray_o'].shape[0]
    resident = {k: v.to(device) for k, v in host.items()}
    eps, alpha = 8.0 / 255.0, 1.0 / 255.0
    base_fm = [f.clone() for f in featmaps]

    def step(batch):
        loss, g_c, g_f = pgd_hot_step(model, projector, batch, featmaps, N_SAMPLES, N_IMPORTANCE, inv_uniform=True,
                                      det=True, max_rays=a.max_rays, group=group)
        with torch.no_grad():                      # sign-step + projection onto the eps-ball (eval_adv.py:824-839)
            for f, g, b in zip(featmaps, (g_c, g_f), base_fm):
                f.add_(alpha * torch.sign(g))
                torch.minimum(torch.maximum(f, b - eps), b + eps, out=f)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = dict(static)
    batch.update(resident)
    for _ in range(a.warmup):
        step(batch)
    barrier()

    # ---------------- timed region: K steps, inputs resident ----------------
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.LAUNCHES
    _lib.profile_start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    t_wall = time.perf_counter()
    for s, e in ev:
        s.record()
        step(batch)
        e.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    prof = _lib.profile_stop()
    launches = _lib.LAUNCHES - launches0
    clocks = sampler.stop() if sampler else None
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = ev[0][0].elapsed_time(ev[-1][1])

    # ---------------- forward-only full-frame pass (rays/s fwd) ----------------
    fwd_ms = []
    with torch.no_grad():
        for i in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for lo in range(0, R, a.max_rays):
                chunk = dict(batch)
                for k in ('ray_o', 'ray_d', 'rgb'):
                    chunk[k] = batch[k][lo:lo + a.max_rays]
                render_rays(chunk, model, featmaps, projector, N_SAMPLES, inv_uniform=True, N_importance=N_IMPORTANCE, det=True)
            e.record()
            torch.cuda.synchronize()
            fwd_ms.append(s.elapsed_time(e))

    # ---------------- the same frame through the reference-facing render_single_image (render_image.py:21-121): the caller's
    # chunk size is the reference's default 4096; every output of the frame (rgb, depth, weights, alpha, z_vals, mask of both
    # levels, ~230 MB) is copied to the host, as the reference's callers expect ----------------
    rsi_ms = None
    try:
        from nerfool_b200.render_image import render_single_image
        sampler = types.SimpleNamespace(H=H, W=W)
        rb = dict(batch)
        with torch.no_grad():
            render_single_image(sampler, rb, model, projector, 4096, N_SAMPLES, inv_uniform=True, N_importance=N_IMPORTANCE, det=True,
                                featmaps=featmaps)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                render_single_image(sampler, rb, model, projector, 4096, N_SAMPLES, inv_uniform=True, N_importance=N_IMPORTANCE,
                                    det=True, featmaps=featmaps)
            torch.cuda.synchronize()
            rsi_ms = (time.perf_counter() - t0) * 500
    except Exception as ex:
        rsi_ms = f'{type(ex).__name__}: {ex}'[:160]

    # ---------------- e2e: host buffers in, loss out, inside the timed region ----------------
    barrier()
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    for s, e in e2e_ev:
        s.record()
        b2 = dict(static)
        for k, v in host.items():
            b2[k] = v.to(device, non_blocking=True)
        loss = step(b2)
        loss_host = loss.to('cpu', non_blocking=False)      # device -> host read of the step's result
        e.record()
    barrier()
    e2e_total_ms = e2e_ev[0][0].elapsed_time(e2e_ev[-1][1])

    # ---------------- the same step in plain-bf16 tensor-core mode (reported beside the headline) ----------------
    bf16_ms = None
    if not a.no_bf16:
        saved_prec = _lib.get_precision()
        _lib.set_precision('bf16')
        step(batch)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(2):
            step(batch)
        s1.record()
        barrier()
        bf16_ms = s0.elapsed_time(s1) / 2
        _lib.set_precision(saved_prec)

    # ---------------- PGD iterations/s at the reference's ray-batch sizes (N_rand, config.py:55 default 512) ----------------
    nrand = {}
    if world == 1 and not a.no_nrand:
        gen = torch.Generator(device='cpu').manual_seed(3)
        for n in (512, 4096, 32768):
            sel = torch.randperm(R, generator=gen)[:n].sort().values.to(device)
            nb = dict(static)
            for k in ('ray_o', 'ray_d', 'rgb'):
                nb[k] = resident[k][sel].contiguous()
            for _ in range(3):
                step(nb)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(10):
                step(nb)
            s1.record()
            torch.cuda.synchronize()
            nrand[str(n)] = {'ms_per_iter': s0.elapsed_time(s1) / 10, 'iters_per_s': 1e4 / s0.elapsed_time(s1)}
            # the same step captured in a CUDA graph (attack.GraphedPGDStep): launch / host overhead removed
            try:
                from nerfool_b200.attack import GraphedPGDStep
                gstep = GraphedPGDStep(model, projector, nb, featmaps, N_SAMPLES, N_IMPORTANCE, inv_uniform=True, det=True,
                                       max_rays=a.max_rays)
                for _ in range(3):
                    gstep(nb['ray_o'], nb['ray_d'], nb['rgb'], featmaps)
                torch.cuda.synchronize()
                s0.record()
                for _ in range(20):
                    gstep(nb['ray_o'], nb['ray_d'], nb['rgb'], featmaps)
                s1.record()
                torch.cuda.synchronize()
                nrand[str(n)]['graph_ms_per_iter'] = s0.elapsed_time(s1) / 20
                nrand[str(n)]['graph_iters_per_s'] = 2e4 / s0.elapsed_time(s1)
                del gstep
            except Exception as ex:
                nrand[str(n)]['graph_error'] = f'{type(ex).__name__}: {ex}'[:160]

    # ---------------- one TRAINING step (train.py:317-327): fwd + loss + backward incl. every IBRNet parameter gradient ----------------
    train = None
    if world == 1 and not a.no_nrand:
        try:
            from nerfool_b200.attack import rgb_loss
            gen = torch.Generator(device='cpu').manual_seed(4)
            sel = torch.randperm(R, generator=gen)[:4096].sort().values.to(device)
            tb = dict(static)
            for k in ('ray_o', 'ray_d', 'rgb'):
                tb[k] = resident[k][sel].contiguous()
            model.net_coarse.train(); model.net_fine.train()
            tfm = tuple(f.detach().clone().requires_grad_(True) for f in featmaps)

            def train_step():
                for n in (model.net_coarse, model.net_fine):
                    n.zero_grad(set_to_none=True)
                out = render_rays(tb, model, tfm, projector, N_SAMPLES, inv_uniform=True, N_importance=N_IMPORTANCE, det=True)
                rgb_loss(out, tb['rgb']).backward()
            for _ in range(3):
                train_step()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(10):
                train_step()
            s1.record()
            torch.cuda.synchronize()
            train = {'rays': 4096, 'ms_per_step': s0.elapsed_time(s1) / 10, 'rays_per_s': 4096e4 / s0.elapsed_time(s1),
                     'what': 'render_rays fwd + masked MSE + backward to the feature maps AND all 2 x 20,136 IBRNet parameters '
                             '(weight gradients as tcgen05 GEMMs over the row index, bf16 operands, fp32 accumulate)'}
        except Exception as ex:
            train = {'error': f'{type(ex).__name__}: {ex}'[:200]}
        finally:
            model.net_coarse.eval(); model.net_fine.eval()

    # max over ranks
    t = torch.tensor([total_ms, e2e_total_ms, statistics.median(fwd_ms), bf16_ms or 0.0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_total_ms, fwd_med, bf16_max = t.tolist()
    bf16_ms = bf16_max if bf16_ms is not None else None

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        tensor_peak = load_tensor_peak()
        traffic_tab = load_traffic_table()
        ktot = {k: sum(v) for k, v in prof.items()}
        kshare = {k: v / sum(ktot.values()) for k, v in ktot.items()}
        # units processed per step by the per-row (view stage) and per-sample (ray stage) kernels
        samples_per_step = R * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE)
        rows_per_step = samples_per_step * a.views
        # ALGORITHMIC bytes per unit (SURVEY.md 8d / DESIGN.md 3): gather 560 B and scatter 512 B per (sample, view) row;
        # the view backward is charged the gather it replaces by reading the activation stash (which actually moves
        # 768 B/row); ray stage: the 288-byte interface row in (fwd) / in + out (bwd) per sample
        alg = {'nfb_ibrnet_view_fwd': (GATHER_B, rows_per_step), 'nfb_ibrnet_view_bwd': (GATHER_B + SCATTER_B, rows_per_step),
               'nfb_ibrnet_ray_fwd': (288 + 16, samples_per_step), 'nfb_ibrnet_ray_bwd': (288 * 2 + 16, samples_per_step)}
        # dense MACs per unit (SURVEY.md 8a FLOP model): 13,256 per row in the view stage, 6,480 + 32 S per sample in the
        # ray stage; the data-gradient is ~1x the forward (dgrad only; the view backward reads the stash, no recompute)
        ray_macs = R * (N_SAMPLES * (6480 + 32 * N_SAMPLES) + (N_SAMPLES + N_IMPORTANCE) * (6480 + 32 * (N_SAMPLES + N_IMPORTANCE)))
        macs = {'nfb_ibrnet_view_fwd': 13256 * rows_per_step, 'nfb_ibrnet_view_bwd': 13256 * rows_per_step,
                'nfb_ibrnet_ray_fwd': ray_macs, 'nfb_ibrnet_ray_bwd': 2 * ray_macs}
        per_kernel = {}
        for k, (bpu, units) in alg.items():
            if k not in prof:
                continue
            n_l = len(prof[k])
            avg_ms = ktot[k] / n_l
            per_launch_units = units * a.steps / n_l
            ach = bpu * per_launch_units / (avg_ms * 1e-3) / 1e9
            tfl = 2 * macs[k] * a.steps / n_l / (avg_ms * 1e-3) / 1e12
            tr = None
            for tab, key in (('dram_bytes_per_row', 'nfb_ibrnet_view'), ('dram_bytes_per_sample', 'nfb_ibrnet_ray')):
                if k.startswith(key) and traffic_tab and k in traffic_tab.get(tab, {}):
                    tr = traffic_tab[tab][k] * per_launch_units
            per_kernel[k] = {'ms_per_step': ktot[k] / a.steps, 'launches_per_step': n_l / a.steps, 'avg_launch_ms': avg_ms,
                             'algorithmic_bytes_per_unit': bpu, 'units_per_launch': per_launch_units,
                             'achieved_GBps': ach, 'hbm_frac': ach / hbm_peak,
                             'dense_TFLOPs': tfl, 'tensor_frac': tfl / tensor_peak if tensor_peak else None,
                             'traffic_bytes_per_launch': tr}
        # the small HBM-bound kernels of the path: compositing and the importance sampler
        s_c, s_f = N_SAMPLES, N_SAMPLES + N_IMPORTANCE
        small = {'nfb_composite_fwd': (29, samples_per_step),        # raw 16 + z 4 + pixel mask 1 in, weights 4 + alpha 4 out per sample
                 'nfb_composite_bwd': (36, samples_per_step),        # raw 16 + z 4 in, d_raw 16 out per sample
                 'nfb_fine_depths': (4 * (2 * s_c + s_f), R)}        # z, weights in, sorted z out per ray
        for k, (bpu, units) in small.items():
            if k in prof:
                n_l = len(prof[k])
                avg_ms = ktot[k] / n_l
                per_launch_units = units * a.steps / n_l
                ach = bpu * per_launch_units / (avg_ms * 1e-3) / 1e9
                per_kernel[k] = {'ms_per_step': ktot[k] / a.steps, 'launches_per_step': n_l / a.steps, 'avg_launch_ms': avg_ms,
                                 'algorithmic_bytes_per_unit': bpu, 'units_per_launch': per_launch_units,
                                 'achieved_GBps': ach, 'hbm_frac': ach / hbm_peak, 'dense_TFLOPs': None, 'tensor_frac': None,
                                 'traffic_bytes_per_launch': None}
        dom = max(per_kernel, key=lambda k: per_kernel[k]['ms_per_step'])
        d = per_kernel[dom]
        ms_per_step = total_ms / a.steps
        rays_total = R * world
        from nerfool_b200 import _lib as _l
        out = {
            'metric': 'rays/s', 'value': rays_total / (ms_per_step * 1e-3), 'unit': 'rays/s',
            'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'BASELINE configs[{a.config}]: IBRNet PGD hot-path step (render_rays fwd + masked-MSE + '
                                   f'bwd to source feature maps), {H}x{W} target view, all {R} rays per step, '
                                   f'{a.views} source views, {N_SAMPLES} coarse + {N_IMPORTANCE} importance samples, random-init weights',
                       'rays_per_step_per_gpu': R, 'source_views': a.views, 'max_rays_per_launch': a.max_rays,
                       'arithmetic': f'{_l.get_precision()}: fp32 data, dense layers on tcgen05 with bf16 hi+lo split operands '
                                     '(3 MMA passes, fp32 accumulate) = fp32-equivalent results' if _l.get_precision() == 'bf16x3'
                                     else _l.get_precision(),
                       'parallelism': f'one target view per GPU x{world}, 1 NCCL allreduce of d(featmaps)/step' if world > 1 else 'single GPU',
                       'l2': 'per-step working set (per-sample workspaces + activation stash, tens of GB) >> 126 MB L2; '
                             'the 2x6.3 MB feature maps are L2-resident by design'},
            'pgd_iters_per_s': 1e3 / ms_per_step,
            'pgd_iters_per_s_by_n_rand': nrand or None,
            'training_step': train,
            'fwd_rays_per_s': rays_total / (fwd_med * 1e-3),
            'fwd_ms_per_frame': fwd_med,
            'render_single_image_ms_per_frame': rsi_ms,
            'encoder': 'not timed: the ResUNet stays on cuDNN in the reference and is outside this repo (north_star)',
            'wall_s_timed_region': t_wall,
            'step_ms': step_ms,
            'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': d['achieved_GBps'], 'peak': hbm_peak, 'unit': 'GB/s',
                         'frac': d['hbm_frac'], 'traffic': d['traffic_bytes_per_launch'], 'peak_source': peak_src,
                         'avg_launch_ms': d['avg_launch_ms'],
                         'algorithmic_bytes_per_launch': d['algorithmic_bytes_per_unit'] * d['units_per_launch'],
                         'traffic_source': (traffic_tab or {}).get('source'),
                         'note': 'algorithmic bytes = SURVEY.md 8(d) per-unit figure x units of the launch; the feature maps are '
                                 'L2-resident, so the gather part of these bytes is served by L2, not DRAM'},
            'roofline_tensor': {'bound': 'tensor', 'kernel': dom, 'achieved': d['dense_TFLOPs'], 'peak': tensor_peak,
                                'unit': 'TFLOP/s', 'frac': d['tensor_frac'],
                                'note': 'useful dense FLOPs of the reference network (not the 3x of the split passes)'},
            'kernels': per_kernel,
            'kernel_share': {k: round(v, 4) for k, v in sorted(kshare.items(), key=lambda kv: -kv[1])},
            'kernel_ms_per_step': {k: v / a.steps for k, v in ktot.items()},
            'e2e': {'value': rays_total / (e2e_total_ms / a.steps * 1e-3), 'unit': 'rays/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'ms_per_step': e2e_total_ms / a.steps},
            'gpu_launches': launches,
            'clocks': clocks,
            'loss_last': float(loss_host),
        }
        if bf16_ms is not None:
            out['bf16_mode'] = {'ms_per_step': bf16_ms, 'rays_per_s': rays_total / (bf16_ms * 1e-3),
                                'note': 'same step with NFB_PREC_BF16 (single bf16 MMA pass; PSNR-parity mode, tests/test_gpu_parity.py::test_precision_modes)'}
        if world == 1 and not a.no_cpu_baseline:
            out['cpu_baseline'] = cpu_reference(a, sample_rays=a.cpu_rays, steps=1, warmup=1)
            out['torch_eager_b200'] = eager_gpu_reference(a, device)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
def cpu_reference(a, sample_rays, steps, warmup):
    """The CPU port of the reference path (oracle/) on a bounded sample of the same workload: the same step
    (render_rays fwd + masked MSE + backward to the feature maps) on `sample_rays` rays of the view."""
    from oracle import ibrnet_oracle as O
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    torch.set_num_threads(os.cpu_count() or 1)
    scene = make_scene(H, W, a.views, seed=0, kind=SCENE_KIND)
    ids = np.sort(np.random.RandomState(1).choice(H * W, sample_rays, replace=False))
    batch = ray_batch_for(scene, ids)
    pc = O.random_ibrnet_params(N_SAMPLES, 1, sigma_bias=0.3)
    pf = O.random_ibrnet_params(N_SAMPLES + N_IMPORTANCE, 2, sigma_bias=0.3)
    times = []
    for i in range(warmup + steps):
        fm = tuple(f.clone().requires_grad_(True) for f in scene['featmaps'])
        t0 = time.perf_counter()
        out = O.render_rays(batch, pc, pf, fm, N_SAMPLES, inv_uniform=True, n_importance=N_IMPORTANCE, det=True)
        loss = O.attack_loss(out, batch['rgb'])
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {'value': sample_rays / sec, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{sample_rays} random rays of the {H}x{W} view, same step (fwd + loss + bwd to feature maps), '
                      f'{warmup} warm-up + mean of {steps}', 'ms_per_step': sec * 1e3}


def eager_gpu_reference(a, device, rays=4096, steps=3):
    """The reference algorithm as eager PyTorch ON THE SAME B200 (the oracle port moved to CUDA, TF32 off): what a
    user of the reference gets today on this GPU.  Same step (render_rays fwd + masked MSE + backward to the feature
    maps) on a chunk of `rays` rays (the reference's own chunking: autograd keeps ~1 GB per 1k rays alive)."""
    from oracle import ibrnet_oracle as O
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        scene = make_scene(H, W, a.views, seed=0, kind=SCENE_KIND)
        ids = np.sort(np.random.RandomState(1).choice(H * W, rays, replace=False))
        batch = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in ray_batch_for(scene, ids).items()}
        pc = {k: v.to(device) for k, v in O.random_ibrnet_params(N_SAMPLES, 1, sigma_bias=0.3).items()}
        pf = {k: v.to(device) for k, v in O.random_ibrnet_params(N_SAMPLES + N_IMPORTANCE, 2, sigma_bias=0.3).items()}
        ms = []
        for i in range(1 + steps):
            fm = tuple(f.to(device).clone().requires_grad_(True) for f in scene['featmaps'])
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = O.render_rays(batch, pc, pf, fm, N_SAMPLES, inv_uniform=True, n_importance=N_IMPORTANCE, det=True)
            O.attack_loss(out, batch['rgb']).backward()
            e.record()
            torch.cuda.synchronize()
            if i > 0:
                ms.append(s.elapsed_time(e))
            del out, fm
        t = statistics.median(ms)
        return {'value': rays / (t * 1e-3), 'unit': 'rays/s', 'ms_per_chunk': t, 'rays_per_chunk': rays,
                'what': 'eager PyTorch (CUDA, fp32, TF32 off) port of the reference path on the same B200, same step; '
                        'reported baseline, not part of the product path'}
    except Exception as ex:       # never let the baseline break the bench line
        return {'unavailable': f'{type(ex).__name__}: {ex}'[:200]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------
# BASELINE configs[4]: GNT render (forward only) of a 378x504 view, 8 source views, 64 samples, trans_depth 4
# ----------------------------------------------------------------------------------------------------
GNT_DEPTH, GNT_VIEWS, GNT_SAMPLES = 4, 8, 64


def gnt_flops_per_ray(S, V, depth):
    # MAC/ray = S [6336 V + depth (9760 V + 90112 + 128 S) + ceil(depth/2) 16256]
    return 2 * S * (6336 * V + depth * (9760 * V + 90112 + 128 * S) + ((depth + 1) // 2) * 16256)


def gnt_cpu_reference(rays, steps=1, warmup=1):
    from oracle import gnt_oracle as G
    from oracle import ibrnet_oracle as O
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    torch.set_num_threads(os.cpu_count() or 1)
    scene = make_scene(H, W, GNT_VIEWS, seed=0, kind='llff')
    ids = np.sort(np.random.RandomState(1).choice(H * W, rays, replace=False))
    batch = ray_batch_for(scene, ids)
    p = G.random_gnt_params(GNT_DEPTH, 1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            pts, z = O.coarse_depths(batch['ray_o'], batch['ray_d'], batch['depth_range'], GNT_SAMPLES, inv_uniform=True, det=True)
            rf, rd, mk = O.projector_compute(pts, batch['camera'], batch['src_rgbs'], batch['src_cameras'], scene['featmaps'][0])
            G.gnt_forward(p, GNT_DEPTH, rf, rd, mk, pts, batch['ray_d'], ret_alpha=True)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {'value': rays / sec, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{rays} random rays of the {H}x{W} view, GNT forward render (projector + network), {warmup} warm-up + mean of {steps}',
            'ms_per_step': sec * 1e3}


def run_gnt(a):
    """One step = the forward render of ALL rays of one 378x504 target view through gnt.render_rays (coarse depths,
    projection + gather, GNT network with ret_alpha, single_net, N_importance = 0: configs/gnt/gnt_llff.txt)."""
    import torch.distributed as dist
    from nerfool_b200 import _lib
    from nerfool_b200.gnt import GNT, render_rays as gnt_render_rays
    from nerfool_b200.projection import Projector
    from nerfool_b200.synthetic import make_scene, rays_for_view
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: nerfool_b200 has no CPU path')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=device)
    _lib.load()
    scene = make_scene(H, W, GNT_VIEWS, seed=0, kind='llff', n_targets=max(world, 1))
    ray_o, ray_d = rays_for_view(scene['camera'][rank], H, W)
    torch.manual_seed(0)
    net = GNT(types.SimpleNamespace(netwidth=64, trans_depth=GNT_DEPTH), 32, 63, 63, ret_alpha=True).to(device).eval()
    model = types.SimpleNamespace(net_coarse=net, net_fine=None)
    projector = Projector(device)
    host = {'ray_o': ray_o.pin_memory(), 'ray_d': ray_d.pin_memory()}
    static = {'depth_range': scene['depth_range'].to(device), 'camera': scene['camera'][rank:rank + 1].to(device),
              'src_rgbs': scene['src_rgbs'].to(device), 'src_cameras': scene['src_cameras'].to(device)}
    featmaps = [f.to(device).contiguous() for f in scene['featmaps']]
    R = ray_o.shape[0]
    chunk = min(a.max_rays, 32768) if a.max_rays > 0 else 16384
    resident = {k: v.to(device) for k, v in host.items()}

    def step(src):
        acc = torch.zeros((), device=device)
        with torch.no_grad():
            for lo in range(0, R, chunk):
                b = dict(static)
                b['ray_o'], b['ray_d'] = src['ray_o'][lo:lo + chunk], src['ray_d'][lo:lo + chunk]
                out = gnt_render_rays(b, model, featmaps, projector, GNT_SAMPLES, inv_uniform=True, N_importance=0, det=True,
                                      ret_alpha=True, single_net=True)
                acc = acc + out['outputs_coarse']['rgb'].sum()
        return acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step(resident)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.LAUNCHES
    _lib.profile_start()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record()
    for _ in range(a.steps):
        step(resident)
    s1.record()
    barrier()
    prof = _lib.profile_stop()
    launches = _lib.LAUNCHES - launches0
    clocks = sampler.stop() if sampler else None
    total_ms = s0.elapsed_time(s1)
    # e2e: rays from pinned host memory, checksum of the rendered colours read back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        chk = step({k: v.to(device, non_blocking=True) for k, v in host.items()}).to('cpu')
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([total_ms, e2e_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()
    if rank == 0:
        tensor_peak = load_tensor_peak()
        ms_per_step = total_ms / a.steps
        ktot = {k: sum(v) / a.steps for k, v in prof.items()}
        net_ms = ktot.get('nfb_gnt_fwd', 0.0)
        fl = gnt_flops_per_ray(GNT_SAMPLES, GNT_VIEWS, GNT_DEPTH) * R
        ach = fl / (net_ms * 1e-3) / 1e12 if net_ms else None
        # nfb_gnt_fwd launches 3 + depth * 4.5 kernels per call: count them for gpu_launches
        # fp32 form: 4 kernels per layer; tensor-core form: 9 (pre, k/v, view core, post, FFN, qkv, ray core, post, FFN)
        per_layer = 4 if _lib.get_precision() == 'fp32' else 9
        per_call = 3 + GNT_DEPTH * per_layer + (GNT_DEPTH + 1) // 2
        n_calls = len(prof.get('nfb_gnt_fwd', []))
        out = {'metric': 'rays/s', 'value': R * world / (ms_per_step * 1e-3), 'unit': 'rays/s', 'n_gpus': world, 'steps': a.steps,
               'warmup': a.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'f32', 'data': 'synthetic',
               'config': {'workload': f'BASELINE configs[4]: GNT forward render of one {H}x{W} view (all {R} rays), {GNT_VIEWS} source views, '
                                      f'{GNT_SAMPLES} samples, trans_depth {GNT_DEPTH}, netwidth 64, ret_alpha, single_net, N_importance 0, '
                                      'random-init weights', 'rays_per_step_per_gpu': R, 'source_views': GNT_VIEWS,
                          'max_rays_per_launch': chunk,
                          'arithmetic': ('fp32 CUDA-core kernels' if _lib.get_precision() == 'fp32' else
                                         f'{_lib.get_precision()}: every 64-wide linear layer on tcgen05 (bf16 hi+lo split operands, fp32 accumulate '
                                         '= fp32-equivalent), attention cores / positional q_fc on the CUDA cores'),
                          'parallelism': f'one target view per GPU x{world}, no collective (render)' if world > 1 else 'single GPU',
                          'l2': 'per-chunk working set (projected view features, ~2 GB) >> 126 MB L2'},
               'roofline': {'bound': 'tensor', 'kernel': 'nfb_gnt_fwd (all kernels of the GNT network, one C-ABI call)',
                            'achieved': ach, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': (ach / tensor_peak) if ach and tensor_peak else None,
                            'traffic': None, 'avg_launch_ms': net_ms / max(n_calls / a.steps, 1),
                            'note': 'algorithmic FLOPs = SURVEY.md Appendix B formula (101 MFLOP/ray at depth 4, S 64, V 8), useful FLOPs only '
                                    '(not the 3x of the split passes); the unfused linear kernels are HBM-bound (rows in / out per layer)'},
               'kernel_ms_per_step': ktot,
               'e2e': {'value': R * world / (e2e_ms / a.steps * 1e-3), 'unit': 'rays/s',
                       'h2d_bytes_per_step': sum(v.numel() * v.element_size() for v in host.values()), 'd2h_bytes_per_step': 4,
                       'ms_per_step': e2e_ms / a.steps},
               'gpu_launches': launches - n_calls + n_calls * per_call, 'clocks': clocks, 'checksum': float(chk)}
        if world == 1 and not a.no_cpu_baseline:
            out['cpu_baseline'] = gnt_cpu_reference(min(a.cpu_rays, 512))
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    if a.config == 4:
        cb = gnt_cpu_reference(min(a.cpu_rays, 512), steps=a.steps, warmup=1)
        print(json.dumps({'impl': 'reference', 'metric': 'rays/s', 'value': cb['value'], 'unit': 'rays/s', 'n_gpus': world,
                          'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True,
                          'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': f'BASELINE configs[4]: GNT forward render, {H}x{W}, {GNT_VIEWS} source views, '
                                                 f'{GNT_SAMPLES} samples, depth {GNT_DEPTH}; CPU port of the reference on a bounded sample'},
                          'cpu_baseline': cb, 'e2e': {'value': cb['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0,
                                                      'd2h_bytes_per_step': 0}}), flush=True)
        return
    cb = cpu_reference(a, sample_rays=a.cpu_rays, steps=a.steps, warmup=max(1, min(a.warmup, 2)))
    out = {'impl': 'reference', 'metric': 'rays/s', 'value': cb['value'], 'unit': 'rays/s', 'n_gpus': world,
           'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': f'BASELINE configs[{a.config}]: IBRNet PGD hot-path step, {H}x{W} target view, '
                                  f'{a.views} source views, {N_SAMPLES} + {N_IMPORTANCE} samples; CPU port of the reference path on a bounded sample',
                      'rays_per_step': a.cpu_rays, 'source_views': a.views},
           'cpu_baseline': cb,
           'e2e': {'value': cb['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--views', type=int, default=4)
    ap.add_argument('--max-rays', dest='max_rays', type=int, default=0, help='rays per launch (0 = sized from the stash budget)')
    ap.add_argument('--cpu-rays', dest='cpu_rays', type=int, default=2048)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-nrand', dest='no_nrand', action='store_true', help='skip the N_rand = 512/4096/32768 PGD iteration timings')
    ap.add_argument('--no-bf16', dest='no_bf16', action='store_true', help='skip the extra plain-bf16 measurement')
    ap.add_argument('--config', type=int, default=1, choices=[1, 2, 3, 4],
                    help='BASELINE.json configs index: 1 = headline (378x504, 4 views, 64+64); 2 = universal-attack shape '
                         '(378x504, 10 views, one target view per GPU); 3 = NeRF-Synthetic shape (800x800, 10 views, 64+128); '
                         '4 = GNT forward render (378x504, 8 views, 64 samples, depth 4)')
    a = ap.parse_args()
    global H, W, N_SAMPLES, N_IMPORTANCE, SCENE_KIND
    if a.config == 2:
        a.views = 10
    elif a.config == 3:
        H, W, N_IMPORTANCE, SCENE_KIND = 800, 800, 128, 'synthetic'
        a.views = 10
    if a.max_rays <= 0:
        # bound the two activation stashes of a chunk (768 B per (sample, view) row) to ~56 GB of the 180 GB: 65,536-ray chunks for
        # the headline config (measured: 32,768-ray chunks 167.7 ms/step, 65,536: 164.3, 98,304: 163.9)
        per_ray = (2 * N_SAMPLES + N_IMPORTANCE) * a.views * 768 + (2 * N_SAMPLES + N_IMPORTANCE) * 560
        a.max_rays = max(4096, min(65536, 1 << int(np.log2(56e9 / per_ray))))
    a.warmup = max(a.warmup, 3) if a.impl == 'ours' else a.warmup
    if a.impl == 'reference':
        run_reference(a)
    elif a.config == 4:
        run_gnt(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
