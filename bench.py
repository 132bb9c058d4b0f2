#!/usr/bin/env python
"""Benchmark of the NeRFool / IBRNet per-ray hot path (BASELINE.json: "rays/s fwd and PGD attack iters/s,
378x504 view, 10 src views, 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1|2|3|4]

Workload (config.workload), default --config 2 = the shape BASELINE.json's metric is quoted on: a synthetic 378x504
LLFF-shaped scene, 10 source views, 64 coarse + 64 importance samples (BASELINE configs[2]; one target view per GPU).
One *step* is one attack iteration of the hot path over ALL 190,512 rays of the target view: render_rays (coarse + fine) ->
masked MSE -> backward to the two source feature maps -> sign-step on the feature maps (so consecutive steps depend on
each other).  With N > 1 each rank renders its own target view (weak scaling) and ONE NCCL allreduce combines the
feature-map gradients.  --config 1 is the 4-source-view shape of configs[0]/[1]; it is also measured as an extra block
("v4_block") in the default run.

value       rays/s through forward + backward, inputs resident in HBM, CUDA-event timed, max over ranks
e2e         the same step through the public API with the step's rays / target colours copied from pinned host memory
            and the loss read back, inside the timed region
roofline    the dominant kernel (largest share of the step) against the measured HBM peak, algorithmic gather/scatter
            bytes (560 B per (sample, view) row gathered, 512 B scattered); "kernels" carries the same for
            all four IBRNet kernels, with the dense-FLOP rate against the measured bf16 tensor peak (roofline_tensor)
encoder / pgd_full_iteration
            the reference's ResUNet (cuDNN; staged copy of the unmodified reference, oracle/stage_reference.py) timed
            separately, and the COMPLETE PGD iteration of eval_adv.py:290-304,693-728 -- encoder(src + delta) ->
            render_rays -> loss -> backward to delta -> Adam step + StepLR + clamps -- at N_rand = 512 / 4096 / 32768 /
            all rays, next to the hot-path-only numbers
strong_scaling (N > 1)
            ONE view's 190,512 rays sharded over the N ranks, encoder sharded over the source views, global loss
            normaliser, one allreduce: the full iteration (attack.delta_gradient_step)
cpu_baseline / --impl reference
            the reference's own render_rays (unmodified modules from the staged copy, kind "reference"; the oracle port
            when no staged copy exists) on a bounded ray sample with all host threads, median of >= 3.
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
REF_STAGED = os.path.join(REPO, 'baseline', '_ref')   # byte-for-byte staged copy of the unmodified reference (git-ignored)


def workload_config(a, world=1):
    """config of the JSON line -- built identically by our arm and by --impl reference."""
    R = H * W
    return {'workload': f'BASELINE configs[{a.config}]: IBRNet PGD hot-path step (render_rays fwd + masked-MSE + bwd to source feature '
                        f'maps), {H}x{W} target view, all {R} rays per step, {a.views} source views, {N_SAMPLES} coarse + '
                        f'{N_IMPORTANCE} importance samples, random-init weights',
            'rays_per_step_per_gpu': R, 'source_views': a.views, 'image': [H, W], 'samples': [N_SAMPLES, N_IMPORTANCE]}


def reference_root():
    """Where the UNMODIFIED reference can be imported from at run time (bench.py never reads /root/reference: it does not
    exist on the GPU box; the staged copy travels with the repo snapshot)."""
    return REF_STAGED if os.path.isdir(os.path.join(REF_STAGED, 'ibrnet')) else None


def load_reference_resunet():
    """ibrnet/feature_network.py:ResUNet of the staged reference, loaded under a private module name (the encoder stays on
    cuDNN and is out of scope for this repo -- north star -- but is part of every real PGD iteration)."""
    root = reference_root()
    if root is None:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location('nfb_ref_feature_network', os.path.join(root, 'ibrnet', 'feature_network.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ResUNet


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
    for name in ('traffic_r02.json', 'traffic_r01.json'):
        path = os.path.join(REPO, 'profiles', name)
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


def _event_time(fn, iters, sync=True):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(iters):
        fn()
    s1.record()
    torch.cuda.synchronize()
    return s0.elapsed_time(s1) / iters


def make_hot_step(model, projector, featmaps, max_rays, group):
    """The bench step: pgd_hot_step + sign-step on the feature maps + projection onto the eps-ball (eval_adv.py:711-728 form)."""
    from nerfool_b200.attack import pgd_hot_step
    eps, alpha = 8.0 / 255.0, 1.0 / 255.0
    base_fm = [f.clone() for f in featmaps]

    def step(batch):
        loss, g_c, g_f = pgd_hot_step(model, projector, batch, featmaps, N_SAMPLES, N_IMPORTANCE, inv_uniform=True,
                                      det=True, max_rays=max_rays, group=group)
        with torch.no_grad():
            for f, g, b in zip(featmaps, (g_c, g_f), base_fm):
                f.add_(alpha * torch.sign(g))
                torch.minimum(torch.maximum(f, b - eps), b + eps, out=f)
        return loss
    return step


def chunk_rays_for(views):
    """Bound the two activation stashes of a chunk to ~56 GB of the 180 GB."""
    per_ray = (2 * N_SAMPLES + N_IMPORTANCE) * views * _stash_row_bytes() + (2 * N_SAMPLES + N_IMPORTANCE) * 560
    return max(4096, min(65536, 1 << int(np.log2(56e9 / per_ray))))


def _stash_row_bytes():
    try:
        from nerfool_b200 import _lib
        return max(64, int(_lib.load().nfb_view_stash_bytes(128 * 1024, 4)) // (128 * 1024 * 4))
    except Exception:
        return 768


def encoder_block(device, views):
    """The reference's ResUNet on cuDNN, fwd and fwd+bwd (to the input image = what carries d featmaps on to delta.grad)."""
    ResUNet = load_reference_resunet()
    if ResUNet is None:
        return None, {'unavailable': 'no staged reference (python oracle/stage_reference.py in the build container)'}
    torch.manual_seed(0)
    enc = ResUNet(coarse_out_ch=32, fine_out_ch=32, coarse_only=False).to(device).eval()
    x = torch.rand(views, 3, H, W, device=device)
    with torch.no_grad():
        for _ in range(3):
            enc(x)
        fwd = _event_time(lambda: enc(x), 5)
    xg = x.clone().requires_grad_(True)

    def fb():
        c, f = enc(xg)
        torch.autograd.backward([c, f], [torch.ones_like(c), torch.ones_like(f)])
        xg.grad = None
    for _ in range(2):
        fb()
    fwdbwd = _event_time(fb, 5)
    return enc, {'what': f'reference ResUNet (ibrnet/feature_network.py, staged unmodified copy) on {views}x3x{H}x{W}, cuDNN, torch default '
                         f'precision flags (cudnn.allow_tf32={torch.backends.cudnn.allow_tf32})',
                 'fwd_ms': fwd, 'fwd_bwd_ms': fwdbwd}


def pgd_full_iteration_block(device, enc, model, projector, static, resident, R, max_rays):
    """COMPLETE PGD iterations (eval_adv.py:290-304 + :693-728) through attack.PGDAttack: encoder(src + delta) -> render_rays on
    N_rand rays (clean colours) -> masked MSE -> backward through the renderer and the encoder to delta -> Adam(-grad) + StepLR
    + clamps."""
    from nerfool_b200.attack import PGDAttack
    out = {}
    gen = torch.Generator(device='cpu').manual_seed(7)
    for n in (512, 4096, 32768, R):
        sel = torch.arange(R) if n == R else torch.randperm(R, generator=gen)[:n].sort().values
        sel = sel.to(device)
        tb = {'camera': static['camera'], 'depth_range': static['depth_range']}
        for k in ('ray_o', 'ray_d', 'rgb'):
            tb[k] = resident[k][sel].contiguous()
        atk = PGDAttack(lambda x: enc(x), model, projector, static, N_SAMPLES, N_IMPORTANCE, epsilon=8.0, use_adam=True, adam_lr=1e-3,
                        lr_step_size=100, lr_gamma=0.5, inv_uniform=True, det=True, max_rays=max_rays,
                        generator=torch.Generator().manual_seed(11))
        for _ in range(2):
            atk.step(tb)
        torch.cuda.synchronize()
        iters = 3 if n == R else 8
        ms = _event_time(lambda: atk.step(tb), iters)
        out[str(n)] = {'ms_per_iter': ms, 'iters_per_s': 1e3 / ms}
        del atk
    return out


def measure_core(a, device, rank, world, group, views, steps, warmup, extras):
    """Build the workload for `views` source views and time the hot-path step.  Returns a dict of raw measurements."""
    from nerfool_b200 import _lib
    from nerfool_b200.render_ray import render_rays
    import torch.distributed as dist
    local = device.index
    scene, model, projector, host, static, featmaps = build_workload(device, rank, views)
    R = host['ray_o'].shape[0]
    max_rays = a.max_rays if a.max_rays > 0 else chunk_rays_for(views)
    resident = {k: v.to(device) for k, v in host.items()}
    step = make_hot_step(model, projector, featmaps, max_rays, group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = dict(static)
    batch.update(resident)
    for _ in range(warmup):
        step(batch)
    barrier()
    m = {'R': R, 'max_rays': max_rays, 'views': views}
    # ---------------- timed region: K steps, inputs resident ----------------
    sampler = ClockSampler(local) if (rank == 0 and extras) else None
    launches0 = _lib.LAUNCHES
    _lib.profile_start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter()
    for s, e in ev:
        s.record()
        step(batch)
        e.record()
    barrier()
    m['t_wall'] = time.perf_counter() - t_wall
    m['prof'] = _lib.profile_stop()
    m['launches'] = _lib.LAUNCHES - launches0
    m['clocks'] = sampler.stop() if sampler else None
    m['step_ms'] = [s.elapsed_time(e) for s, e in ev]
    m['total_ms'] = ev[0][0].elapsed_time(ev[-1][1])

    # ---------------- the same K steps WITHOUT the per-call event pairs of the profiler ----------------
    barrier()
    m['unprofiled_ms_per_step'] = _event_time(lambda: step(batch), max(2, steps // 2))
    barrier()

    # ---------------- forward-only full-frame pass (rays/s fwd) ----------------
    fwd_ms = []
    with torch.no_grad():
        for i in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for lo in range(0, R, max_rays):
                chunk = dict(batch)
                for k in ('ray_o', 'ray_d', 'rgb'):
                    chunk[k] = batch[k][lo:lo + max_rays]
                render_rays(chunk, model, featmaps, projector, N_SAMPLES, inv_uniform=True, N_importance=N_IMPORTANCE, det=True)
            e.record()
            torch.cuda.synchronize()
            fwd_ms.append(s.elapsed_time(e))
    m['fwd_ms'] = statistics.median(fwd_ms)

    # ---------------- e2e: host buffers in, loss out, inside the timed region ----------------
    barrier()
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    m['h2d'] = sum(v.numel() * v.element_size() for v in host.values())
    for s, e in e2e_ev:
        s.record()
        b2 = dict(static)
        for k, v in host.items():
            b2[k] = v.to(device, non_blocking=True)
        loss = step(b2)
        loss_host = loss.to('cpu', non_blocking=False)      # device -> host read of the step's result
        e.record()
    barrier()
    m['e2e_total_ms'] = e2e_ev[0][0].elapsed_time(e2e_ev[-1][1])
    m['loss_last'] = float(loss_host)
    m['objects'] = (scene, model, projector, host, static, featmaps, resident, step, batch)
    return m


def run_ours(a):
    import torch.distributed as dist
    from nerfool_b200 import _lib
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

    m = measure_core(a, device, rank, world, group, a.views, a.steps, a.warmup, extras=True)
    scene, model, projector, host, static, featmaps, resident, step, batch = m['objects']
    R, max_rays = m['R'], m['max_rays']
    prof, launches, clocks = m['prof'], m['launches'], m['clocks']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- the same frame through the reference-facing render_single_image (render_image.py:21-121): the caller's
    # chunk size is the reference's default 4096; every output of the frame is copied to the host, as the reference's callers expect
    rsi_ms = None
    if world == 1 and not a.no_nrand:
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

    # ---------------- the same step in plain-bf16 tensor-core mode (reported beside the headline) ----------------
    bf16_ms = None
    if not a.no_bf16:
        saved_prec = _lib.get_precision()
        _lib.set_precision('bf16')
        step(batch)
        barrier()
        bf16_ms = _event_time(lambda: step(batch), 2)
        barrier()
        _lib.set_precision(saved_prec)

    # ---------------- PGD iterations/s of the HOT PATH ONLY at the reference's ray-batch sizes (N_rand, config.py:55 default 512)
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
            ms = _event_time(lambda: step(nb), 10)
            nrand[str(n)] = {'ms_per_iter': ms, 'iters_per_s': 1e3 / ms}
            # the same step captured in a CUDA graph (attack.GraphedPGDStep): launch / host overhead removed
            try:
                from nerfool_b200.attack import GraphedPGDStep
                gstep = GraphedPGDStep(model, projector, nb, featmaps, N_SAMPLES, N_IMPORTANCE, inv_uniform=True, det=True,
                                       max_rays=max_rays)
                for _ in range(3):
                    gstep(nb['ray_o'], nb['ray_d'], nb['rgb'], featmaps)
                torch.cuda.synchronize()
                ms = _event_time(lambda: gstep(nb['ray_o'], nb['ray_d'], nb['rgb'], featmaps), 20)
                nrand[str(n)]['graph_ms_per_iter'] = ms
                nrand[str(n)]['graph_iters_per_s'] = 1e3 / ms
                del gstep
            except Exception as ex:
                nrand[str(n)]['graph_error'] = f'{type(ex).__name__}: {ex}'[:160]

    # ---------------- the encoder (cuDNN, reference ResUNet) and the COMPLETE PGD iteration ----------------
    encoder, pgd_full = None, None
    if world == 1 and not a.no_nrand:
        try:
            enc, encoder = encoder_block(device, a.views)
            if enc is not None:
                pgd_full = pgd_full_iteration_block(device, enc, model, projector, static, resident, R, max_rays)
                del enc
        except Exception as ex:
            encoder = encoder or {}
            encoder['error'] = f'{type(ex).__name__}: {ex}'[:300]
        torch.cuda.empty_cache()

    # ---------------- strong scaling (N > 1): ONE view's rays sharded over the ranks, encoder sharded over the source views ----------------
    strong = None
    try:
        strong = strong_scaling_block(a, device, rank, world, group, model, projector, static, resident, R, max_rays)
    except Exception as ex:
        strong = {'error': f'{type(ex).__name__}: {ex}'[:300]}

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
            ms = _event_time(train_step, 10)
            train = {'rays': 4096, 'ms_per_step': ms, 'rays_per_s': 4096e3 / ms,
                     'what': 'render_rays fwd + masked MSE + backward to the feature maps AND all 2 x 20,136 IBRNet parameters '
                             '(weight gradients as tcgen05 GEMMs over the row index, bf16 operands, fp32 accumulate)'}
        except Exception as ex:
            train = {'error': f'{type(ex).__name__}: {ex}'[:200]}
        finally:
            model.net_coarse.eval(); model.net_fine.eval()

    # ---------------- the 4-source-view shape (BASELINE configs[0]/[1]) as an extra block ----------------
    v4 = None
    if world == 1 and not a.no_nrand and a.views != 4 and a.config == 2:
        del scene, model, projector, host, static, featmaps, resident, step, batch
        m['objects'] = None
        torch.cuda.empty_cache()
        try:
            m4 = measure_core(a, device, rank, world, group, 4, max(3, a.steps // 4), 3, extras=False)
            ms4 = m4['total_ms'] / len(m4['step_ms'])
            k4 = {k: sum(v) / len(m4['step_ms']) for k, v in m4['prof'].items()}
            v4 = {'workload': f'BASELINE configs[1]: same step, 4 source views, all {m4["R"]} rays', 'ms_per_step': ms4,
                  'rays_per_s': m4['R'] / (ms4 * 1e-3), 'pgd_iters_per_s': 1e3 / ms4, 'fwd_ms_per_frame': m4['fwd_ms'],
                  'fwd_rays_per_s': m4['R'] / (m4['fwd_ms'] * 1e-3), 'e2e_rays_per_s': m4['R'] / (m4['e2e_total_ms'] / len(m4['step_ms']) * 1e-3),
                  'kernel_ms_per_step': k4, 'max_rays_per_launch': m4['max_rays']}
            m4['objects'] = None
        except Exception as ex:
            v4 = {'error': f'{type(ex).__name__}: {ex}'[:200]}

    # max over ranks
    total_ms, e2e_total_ms, fwd_med = m['total_ms'], m['e2e_total_ms'], m['fwd_ms']
    t = torch.tensor([total_ms, e2e_total_ms, fwd_med, bf16_ms or 0.0, m['unprofiled_ms_per_step']], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_total_ms, fwd_med, bf16_max, unprof = t.tolist()
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
        # the view backward is charged the gather it replaces by reading the activation stash; ray stage: the 288-byte
        # interface row in (fwd) / in + out (bwd) per sample
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
        cfg = workload_config(a, world)
        cfg.update({'max_rays_per_launch': max_rays,
                    'arithmetic': f'{_lib.get_precision()}: fp32 data, dense layers on tcgen05 with bf16 hi+lo split operands '
                                  '(3 MMA passes, fp32 accumulate) = fp32-equivalent results' if _lib.get_precision() == 'bf16x3'
                                  else _lib.get_precision(),
                    'parallelism': f'one target view per GPU x{world}, 1 NCCL allreduce of d(featmaps)/step' if world > 1 else 'single GPU',
                    'l2': 'per-step working set (per-sample workspaces + activation stash, tens of GB) >> 126 MB L2; '
                          'the 2 x 15.7 MB feature maps are L2-resident by design'})
        out = {
            'metric': 'rays/s', 'value': rays_total / (ms_per_step * 1e-3), 'unit': 'rays/s',
            'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': cfg,
            'pgd_iters_per_s': 1e3 / ms_per_step,
            'pgd_iters_per_s_note': 'hot path only (render_rays fwd + loss + bwd to the feature maps), all rays of the view per '
                                    'iteration; the complete iteration incl. the cuDNN encoder and the optimiser is pgd_full_iteration',
            'pgd_iters_per_s_by_n_rand_hot_path_only': nrand or None,
            'encoder': encoder,
            'pgd_full_iteration': pgd_full,
            'strong_scaling': strong,
            'training_step': train,
            'v4_block': v4,
            'fwd_rays_per_s': rays_total / (fwd_med * 1e-3),
            'fwd_ms_per_frame': fwd_med,
            'render_single_image_ms_per_frame': rsi_ms,
            'wall_s_timed_region': m['t_wall'],
            'ms_per_step_without_event_pairs': unprof,
            'step_ms': m['step_ms'],
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
                    'h2d_bytes_per_step': m['h2d'], 'd2h_bytes_per_step': 4, 'ms_per_step': e2e_total_ms / a.steps},
            'gpu_launches': launches,
            'clocks': clocks,
            'loss_last': m['loss_last'],
        }
        if bf16_ms is not None:
            out['bf16_mode'] = {'ms_per_step': bf16_ms, 'rays_per_s': rays_total / (bf16_ms * 1e-3),
                                'note': 'same step with NFB_PREC_BF16 (single bf16 MMA pass; PSNR-parity mode, tests/test_gpu_parity.py::test_precision_modes)'}
        if world == 1 and not a.no_cpu_baseline:
            out['cpu_baseline'] = cpu_reference(a, sample_rays=a.cpu_rays, steps=3, warmup=1)
            out['torch_eager_b200'] = eager_gpu_reference(a, device)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_scaling_block(a, device, rank, world, group, model, projector, static, resident, R, max_rays):
    """SURVEY.md 8(d) "fixed global batch": ONE target view's R rays sharded over the ranks (contiguous slices), the global
    mask count as loss normaliser, the reference encoder sharded over the source views, one packed allreduce, delta-gradient
    slices all-gathered (attack.delta_gradient_step).  Reported at every N (N = 1 is the same program without collectives)."""
    import torch.distributed as dist
    from nerfool_b200.attack import delta_gradient_step, shard_slice
    ResUNet = load_reference_resunet()
    if ResUNet is None or a.no_nrand:
        return None
    torch.manual_seed(0)
    enc = ResUNet(coarse_out_ch=32, fine_out_ch=32, coarse_only=False).to(device).eval()
    if world > 1:                       # every rank works on RANK 0's target view
        for k in ('ray_o', 'ray_d', 'rgb'):
            dist.broadcast(resident[k], src=0)
        cam = static['camera'].clone()
        dist.broadcast(cam, src=0)
    else:
        cam = static['camera']
    lo, hi = shard_slice(R, rank, world)
    shard = {'camera': cam, 'depth_range': static['depth_range'], 'src_rgbs': static['src_rgbs'], 'src_cameras': static['src_cameras']}
    for k in ('ray_o', 'ray_d', 'rgb'):
        shard[k] = resident[k][lo:hi].contiguous()
    delta = ((torch.rand(static['src_rgbs'].shape, generator=torch.Generator().manual_seed(5)) * 2 - 1) * (8. / 255.)).to(device)

    def it():
        return delta_gradient_step(lambda x: enc(x), model, projector, shard, delta, N_SAMPLES, N_IMPORTANCE, inv_uniform=True, det=True,
                                   max_rays=max_rays, group=group, global_norm=True, shard_encoder=True)
    for _ in range(2):
        it()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    iters = 4
    ms = _event_time(it, iters)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    del enc
    torch.cuda.empty_cache()
    return {'what': f'one {H}x{W} view ({R} rays) sharded over {world} rank(s), encoder sharded over the {a.views} source views, full '
                    'iteration to d delta (encoder fwd + render fwd/bwd + encoder bwd + collectives)', 'ms_per_iter': ms,
            'iters_per_s': 1e3 / ms, 'rays_per_s': R / (ms * 1e-3), 'n_gpus': world}


# ----------------------------------------------------------------------------------------------------
def _reference_modules():
    """The UNMODIFIED reference modules of the path, imported from the staged copy (None when it is absent)."""
    root = reference_root()
    if root is None:
        return None
    if root not in sys.path:
        sys.path.append(root)
    from ibrnet.projection import Projector as RefProjector
    from ibrnet.mlp_network import IBRNet as RefIBRNet
    from ibrnet import render_ray as ref_rr
    assert os.path.samefile(os.path.dirname(ref_rr.__file__), os.path.join(root, 'ibrnet')), 'ibrnet resolved to something other than the staged reference'
    return RefProjector, RefIBRNet, ref_rr


def cpu_reference(a, sample_rays, steps, warmup):
    """The reference's own CPU implementation of the path on a bounded sample of the same workload: the same step
    (render_rays fwd + masked MSE + backward to the feature maps) on `sample_rays` rays of the view, all host threads.
    kind "reference": the unmodified ibrnet.{projection, mlp_network, render_ray} from the staged copy; kind "port": the
    oracle restatement (only when no staged copy exists)."""
    from nerfool_b200.synthetic import make_scene, ray_batch_for
    torch.set_num_threads(os.cpu_count() or 1)
    scene = make_scene(H, W, a.views, seed=0, kind=SCENE_KIND)
    ids = np.sort(np.random.RandomState(1).choice(H * W, sample_rays, replace=False))
    batch = ray_batch_for(scene, ids)
    mods = _reference_modules()
    if mods is not None:
        RefProjector, RefIBRNet, ref_rr = mods
        args = types.SimpleNamespace(anti_alias_pooling=1, local_rank=0)
        torch.manual_seed(0)
        nc, nf = RefIBRNet(args, in_feat_ch=32, n_samples=N_SAMPLES).eval(), RefIBRNet(args, in_feat_ch=32, n_samples=N_SAMPLES + N_IMPORTANCE).eval()
        with torch.no_grad():
            for n in (nc, nf):
                n.out_geometry_fc[2].bias += 0.3
        model, proj = types.SimpleNamespace(net_coarse=nc, net_fine=nf), RefProjector(device='cpu')

        def one(fm):
            ret = ref_rr.render_rays(batch, model, fm, proj, N_samples=N_SAMPLES, inv_uniform=True, N_importance=N_IMPORTANCE, det=True)
            loss = 0.
            for lvl in ('outputs_coarse', 'outputs_fine'):
                mk = ret[lvl]['mask'].float()
                loss = loss + torch.sum((ret[lvl]['rgb'] - batch['rgb']) ** 2 * mk[:, None]) / (torch.sum(mk) * 3 + 1e-6)   # utils.img2mse
            loss.backward()
        kind = 'reference'
    else:
        from oracle import ibrnet_oracle as O
        pc = O.random_ibrnet_params(N_SAMPLES, 1, sigma_bias=0.3)
        pf = O.random_ibrnet_params(N_SAMPLES + N_IMPORTANCE, 2, sigma_bias=0.3)

        def one(fm):
            out = O.render_rays(batch, pc, pf, fm, N_SAMPLES, inv_uniform=True, n_importance=N_IMPORTANCE, det=True)
            O.attack_loss(out, batch['rgb']).backward()
        kind = 'port'
    times = []
    for i in range(warmup + steps):
        fm = tuple(f.clone().requires_grad_(True) for f in scene['featmaps'])
        t0 = time.perf_counter()
        one(fm)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = statistics.median(times)
    return {'value': sample_rays / sec, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': kind,
            'sample': f'{sample_rays} random rays of the {H}x{W} view ({a.views} source views), same step (fwd + loss + bwd to feature maps), '
                      f'{warmup} warm-up + median of {steps}' + ('; unmodified reference modules (staged copy)' if kind == 'reference'
                                                                   else '; oracle port (no staged reference present)'),
            'ms_per_step': sec * 1e3, 'step_s': times}


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
    # GNT attack step (eval/gnt/eval_adv.py:282-545 below the encoder): render N_rand rays -> MSE -> backward to the feature maps
    attack = None
    if world == 1 and not a.no_nrand:
        attack = {}
        for nr in (512, 4096):
            idx = torch.arange(0, R, max(R // nr, 1), device=device)[:nr]
            b = dict(static)
            b['ray_o'], b['ray_d'] = resident['ray_o'][idx].contiguous(), resident['ray_d'][idx].contiguous()
            tgt = torch.rand(idx.numel(), 3, device=device)
            fm = [f.clone().requires_grad_(True) for f in featmaps]

            def astep():
                out = gnt_render_rays(b, model, fm, projector, GNT_SAMPLES, inv_uniform=True, N_importance=0, det=True,
                                      ret_alpha=True, single_net=True)
                loss = ((out['outputs_coarse']['rgb'] - tgt) ** 2).mean()
                return torch.autograd.grad(loss, fm[0])[0]
            astep(); astep()
            torch.cuda.synchronize()
            ms = _event_time(astep, 5)
            _lib.profile_start()
            astep()
            torch.cuda.synchronize()
            kp = {k: sum(v) for k, v in _lib.profile_stop().items()}
            attack[f'N_rand_{nr}'] = {'ms_per_iter': ms, 'iters_per_s': 1e3 / ms, 'rays_per_s': idx.numel() / (ms * 1e-3), 'kernel_ms': kp}
        attack['note'] = ('hot path only (no encoder): gnt.render_rays -> MSE -> d featmaps; nfb_gnt_bwd = fp32 checkpointing forward + '
                          'block-wise reverse sweep on the CUDA cores (ReLU masks of the fp32 forward), scatter by nfb_project_gather_bwd')
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
               'kernel_ms_per_step': ktot, 'attack_step': attack,
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
           'config': workload_config(a, world),
           'cpu_baseline': cb,
           'e2e': {'value': cb['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--views', type=int, default=0, help='source views (0 = what --config names)')
    ap.add_argument('--max-rays', dest='max_rays', type=int, default=0, help='rays per launch (0 = sized from the stash budget)')
    ap.add_argument('--cpu-rays', dest='cpu_rays', type=int, default=0, help='rays of the bounded CPU sample per step (0 = 512 at 10 views, 2048 at 4)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-nrand', dest='no_nrand', action='store_true', help='skip the N_rand = 512/4096/32768 PGD iteration timings')
    ap.add_argument('--no-bf16', dest='no_bf16', action='store_true', help='skip the extra plain-bf16 measurement')
    ap.add_argument('--config', type=int, default=2, choices=[1, 2, 3, 4],
                    help='BASELINE.json configs index: 2 (default) = the shape the metric is quoted on (378x504, 10 source views, '
                         '64+64, one target view per GPU); 1 = the 4-source-view shape of configs[0]/[1]; 3 = NeRF-Synthetic '
                         'shape (800x800, 10 views, 64+128); 4 = GNT forward render (378x504, 8 views, 64 samples, depth 4)')
    a = ap.parse_args()
    global H, W, N_SAMPLES, N_IMPORTANCE, SCENE_KIND
    if a.config == 3:
        H, W, N_IMPORTANCE, SCENE_KIND = 800, 800, 128, 'synthetic'
    if a.views <= 0:
        a.views = {1: 4, 2: 10, 3: 10, 4: GNT_VIEWS}[a.config]
    if a.cpu_rays <= 0:
        a.cpu_rays = 2048 if a.views <= 4 else 512
    a.warmup = max(a.warmup, 3) if a.impl == 'ours' else a.warmup
    if a.impl == 'reference':
        run_reference(a)
    elif a.config == 4:
        run_gnt(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
