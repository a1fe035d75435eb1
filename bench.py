#!/usr/bin/env python
"""Benchmark of the LCGP emulator-fitting hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4_rep] [--impl ours|reference]

A "step" is one objective + analytic-gradient evaluation over all q latents at the `init_params`
point of the named synthetic configuration (default: BASELINE.json config 4, n=8000, d=10, p=2000,
q=32, replicated data).  For N > 1 launch under torchrun; the latents are sharded over the ranks and
each step ends with the all-reduce of the objective / gradient, so `value` is whole-job evals/s
(total work fixed as N grows: "strong" scaling).

  value : K device-resident steps (parameters already in HBM, no host copies) timed with CUDA events
  e2e   : K steps through LCGP.loss_and_grad() -- host parameters in, host objective + gradient out
  roofline / stages : per-stage CUDA-event times recorded inside the C-ABI call on its stream
  cpu_baseline : the torch-float64 oracle timed on this box's host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='cfg4_rep')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fit', action='store_true', help='skip the end-to-end fit() wall-time measurement')
    ap.add_argument('--no-named-configs', action='store_true',
                    help='skip the evals/s of BASELINE configs 1, 2, 3, 5 at their named shapes (N = 1 only, ~1 min with the CPU port)')
    ap.add_argument('--fit-config', default='cfg3_rep',
                    help="configuration of the fit() wall-time measurement; 'workload' = the bench configuration itself "
                         '(several minutes at config 4 on one GPU); under torchrun (N > 1) the default is the workload, '
                         'fitted sharded over the ranks')
    ap.add_argument('--cpu-sample-latents', type=int, default=1)
    ap.add_argument('--emulators', type=int, default=64)      # cfg5_batch only
    ap.add_argument('--threads', type=int, default=4)         # cfg5_batch only: host threads (streams) per GPU (more only contend for the GIL)
    ap.add_argument('--maxiter', type=int, default=0)         # cfg5_batch only: L-BFGS-B iteration budget per emulator (0 = to convergence)
    ap.add_argument('--engine', default='auto', choices=['auto', 'lockstep', 'threads'])   # cfg5_batch only (lcgp_b200.batched)
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc = None
        self.lines = []
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [s.strip() for s in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'power_w_max': float(max(pw)),
                'samples': len(sm), 'reasons': sorted(reasons)}


def dgemm_peak(device, n=8192, reps=4, sustain_s=4.0):
    """cuBLAS DGEMM throughput on this GPU: the FP64 tensor-pipe denominator.  Returns (burst, sustained) TFLOP/s:
    best of `reps` single 8192^3 products, and the average over back-to-back products for `sustain_s` seconds."""
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    c = torch.empty(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flop = 2.0 * n ** 3
    k = max(4, int(sustain_s / (best * 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    sustained = k * flop / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    return flop / (best * 1e-3) / 1e12, sustained


def build_model(cfg_name):
    from lcgp_b200 import LCGP, synthetic
    x, y, x0, y0, mk = synthetic.make_config(cfg_name)
    # lazy one-off initialisation (CUDA context, cuBLAS handle, NCCL communicator) is not construction time
    dev = torch.device('cuda', torch.cuda.current_device())
    a = torch.ones(256, 256, dtype=torch.float64, device=dev)
    (a @ a).sum().item()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(a)
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    model = LCGP(y=y, x=x, **mk)
    return model, time.time() - t0, (x, y, mk)


def workload_desc(cfg_name, n, d, p, q, submethod, world):
    """`config` of the JSON line -- IDENTICAL in the CUDA arm and the reference arm (arm-specific text lives in the
    top-level `arm` key), so that the driver's same_config comparison holds."""
    npad = (n + 127) // 128 * 128
    return {'workload': f'{cfg_name}: objective+gradient, n={n} unique inputs, d={d}, p={p}, q={q}, '
                        f'submethod={submethod}, init_params point',
            'n': n, 'd': d, 'p': p, 'q': q,
            'l2': 'inputs larger than L2 (%.1f GB of dense n x n factors per evaluation)' % (q * npad ** 2 * 8 / 1e9),
            'parallelism': f'latents sharded over {world} rank(s)'}


def cpu_baseline_sample(x, y, mk, n_latents, q, threads):
    """Oracle (torch float64 CPU, autograd) on `n_latents` of the q latents at full n; evals/s
    extrapolated as 1 / (q/n_latents * t).  Returns (evals_per_s, seconds, description)."""
    from oracle.lcgp_oracle import LCGPOracle
    import psutil
    torch.set_num_threads(threads)
    n_unique = np.unique(x, axis=0).shape[0] if mk['submethod'] == 'rep' else x.shape[0]
    need = (4 * x.shape[1] + 12) * n_unique * n_unique * 8 * n_latents   # autograd keeps ~4 n x n per input dim
    if need > 0.6 * psutil.virtual_memory().available:
        raise MemoryError(f'oracle sample needs ~{need / 1e9:.0f} GB of host memory')
    o = LCGPOracle(y=y, x=x, skip_xnorm=True, **mk)
    lat = list(range(n_latents))
    fn = (lambda: o.neglpost_rep(latents=lat)) if mk['submethod'] == 'rep' else None
    if fn is None:
        fn = lambda: o.neglpost_chol(latents=lat)
    t0 = time.time()
    o.loss_and_grad(fn)
    dt = time.time() - t0
    return o, fn, dt


def fit_wall(cfg_name, with_cpu, cpu_s_per_eval=None, sharded=False):
    """End-to-end fit() wall time (BASELINE metric 'fit wall-s'): constructor excluded, SciPy L-BFGS-B
    defaults.  The reference's fit time is extrapolated as evals x per-eval (labelled as such): per-eval from
    one oracle evaluation on the host cores for configurations where that is affordable, else from the
    bounded-sample CPU baseline of the same workload (`cpu_s_per_eval`).  With sharded=True every rank calls
    this (the latents of the fitted model are sharded over the ranks, one all-reduce per evaluation)."""
    from lcgp_b200 import LCGP, synthetic
    x, y, x0, y0, mk = synthetic.make_config(cfg_name)
    t0 = time.perf_counter()
    m = LCGP(y=y, x=x, shard=sharded, **mk)
    ctor = time.perf_counter() - t0
    m.loss_and_grad()                       # workspace allocation / first-touch outside the timed region
    m.n_evals = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.fit()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out = {'config': cfg_name, 'n': int(m.n), 'q': int(m.q), 'wall_s': wall, 'evals': m.n_evals, 'ctor_s': ctor,
           'final_loss': float(m.loss()), 'optimizer': 'L-BFGS-B (SciPy defaults)',
           'converged': bool(getattr(m.opt_result, 'success', False)), 'iterations': int(getattr(m.opt_result, 'nit', -1))}
    if cpu_s_per_eval is not None:
        out['cpu_port_s_per_eval'] = cpu_s_per_eval
        out['cpu_port_fit_s_extrapolated'] = cpu_s_per_eval * m.n_evals
        out['cpu_note'] = 'per-eval from the bounded-sample cpu_baseline of this run'
    elif with_cpu and int(m.n) * int(m.q) <= 40000:
        try:
            from oracle.lcgp_oracle import LCGPOracle
            torch.set_num_threads(os.cpu_count() or 1)
            o = LCGPOracle(y=y, x=x, skip_xnorm=True, **mk)
            fn = o.neglpost_chol if mk['submethod'] == 'full' else None
            t1 = time.perf_counter()
            o.loss_and_grad(fn)
            per_eval = time.perf_counter() - t1
            out['cpu_port_s_per_eval'] = per_eval
            out['cpu_port_fit_s_extrapolated'] = per_eval * m.n_evals
        except Exception as ex:
            out['cpu_port_s_per_eval'] = f'failed: {ex!r}'
    return out


def named_configs(with_cpu, reps=20):
    """BASELINE.json configs 1, 2, 3, 5 at their NAMED shapes (config 4 is the bench workload itself): objective +
    gradient evaluations per second through LCGP.loss_and_grad() (host parameters in, host gradient out), beside one
    oracle evaluation (torch float64 CPU port, forward + autograd backward) on this box's host cores."""
    from lcgp_b200 import LCGP, synthetic
    cases = []
    x, y, _, _ = synthetic.rep1d_skewed()
    cases += [('cfg1_rep1d_q2', x, y, dict(q=2, submethod='rep')), ('cfg1_rep1d_q3_notebook', x, y, dict(q=3, submethod='rep'))]
    x, y, _ = synthetic.rep3d()
    cases.append(('cfg2_rep3d', x, y, dict(q=3, submethod='rep')))
    for name in ('cfg3_rep', 'cfg3_full', 'cfg5_one'):
        x, y, _, _, mk = synthetic.make_config(name)
        cases.append((name, x, y, mk))
    out = {}
    for name, x, y, mk in cases:
        m = LCGP(y=y, x=x, shard=False, **mk)
        for _ in range(3):
            f, g = m.loss_and_grad()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            m.loss_and_grad()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        ent = {'n': int(m.n), 'd': int(m.d), 'p': int(m.p), 'q': int(m.q), 'submethod': m.submethod,
               'ms_per_eval': dt * 1e3, 'evals_per_s': 1.0 / dt, 'objective': f}
        if with_cpu:
            try:
                from oracle.lcgp_oracle import LCGPOracle
                torch.set_num_threads(os.cpu_count() or 1)
                o = LCGPOracle(y=y, x=x, skip_xnorm=True, **mk)
                fn = o.neglpost_chol if mk['submethod'] == 'full' else None
                o.loss_and_grad(fn)
                k = 1 if int(m.n) >= 1000 else 5
                t1 = time.perf_counter()
                for _ in range(k):
                    fo, _g = o.loss_and_grad(fn)
                per = (time.perf_counter() - t1) / k
                ent.update(cpu_port_s_per_eval=per, cpu_cores=os.cpu_count() or 1, speedup_vs_cpu_port=per / dt,
                           objective_rel_diff_vs_port=abs(f - fo) / abs(fo))
            except Exception as ex:
                ent['cpu_port_s_per_eval'] = f'failed: {ex!r}'
        out[name] = ent
        del m
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    """Reference arm: the oracle port of the reference's CPU path on this box's host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from lcgp_b200 import synthetic
    threads = os.cpu_count() or 1
    x, y, x0, y0, mk = synthetic.make_config(args.config)
    o, fn, first = cpu_baseline_sample(x, y, mk, args.cpu_sample_latents, mk['q'], threads)
    q = int(mk['q'])
    for _ in range(max(args.warmup - 1, 0)):
        o.loss_and_grad(fn)
    t0 = time.time()
    for _ in range(args.steps):
        o.loss_and_grad(fn)
    dt = (time.time() - t0) / args.steps
    per_eval = dt * q / args.cpu_sample_latents
    val = 1.0 / per_eval
    sample = (f'{args.cpu_sample_latents} of {q} latents per step at full n (forward + autograd backward), '
              f'extrapolated x{q // args.cpu_sample_latents}: the reference loop over latents is sequential and every '
              f'latent costs the same (linearity checked once over all latents: profiles/r2_cpu_full_eval_check.txt)')
    line = {'impl': 'reference', 'metric': 'NLL+grad evals/s', 'value': val, 'unit': 'evals/s', 'n_gpus': args.gpus,
            # a step of this arm is the BOUNDED SAMPLE it really executes (ms_per_step = its measured duration, so that
            # steps x ms_per_step is the arm's true timed region); `value` is the whole-workload metric extrapolated from it
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'ms_per_eval_extrapolated': per_eval * 1e3,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_desc(args.config, int(o.n), int(o.d), int(o.p), q, mk['submethod'], args.gpus),
            'arm': f'oracle port of the reference CPU path (oracle/lcgp_oracle.py) on {threads} host threads, rank 0 only',
            # value and ms_per_step are EXTRAPOLATED from the bounded sample each step times
            'extrapolated': True, 'sampled_latents': args.cpu_sample_latents, 'sample_fraction': args.cpu_sample_latents / q,
            'cpu_baseline': {'value': val, 'unit': 'evals/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def run_ours(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (B200); there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))   # N ranks preprocess concurrently on one host
    from lcgp_b200 import _cabi

    model, t_ctor, (x, y, mk) = build_model(args.config)
    eng = model.engine
    q, d, p, n = int(model.q), int(model.d), int(model.p), int(model.n)
    q_loc = len(model._local_idx)
    lLmb, lLmb0, lsig_p, lnug = model.get_param()
    idx = model._local_idx

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def device_step(events=None):
        # N = 1: one lcgp_nll_grad call, parameters and results stay in HBM.
        # N > 1: the same on each rank's latents + the single flat all-reduce (model._evaluate_sharded).
        if world > 1:
            return model._evaluate_sharded(lLmb, lLmb0, lsig_p, lnug, True, events)
        return eng.evaluate_device(lLmb[idx], lLmb0[idx], lnug[idx], lsig_p, True, events)

    # ---- warm-up (also through the public API) ----
    for _ in range(max(args.warmup, 3)):
        chk = device_step()
    if world > 1 and float(chk[-1]) != 0.0:     # last slot of the all-reduced vector: failed latents over all ranks
        raise SystemExit('bench: Cholesky failure reported by the sharded evaluation')
    model.loss_and_grad()                       # (single rank: checks `info` itself)
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()

    # ---- value: device-resident steps, CUDA events on the launching stream ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = int(_cabi.lib().lcgp_launch_count())
    e0.record()
    for s in range(args.steps):
        device_step()
    e1.record()
    launches = int(_cabi.lib().lcgp_launch_count()) - launches0      # counted by the library at each launch site
    barrier()
    t_dev = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)

    # ---- stage breakdown: the same K steps with stage events recorded inside the C-ABI call (this
    #      makes the library join its stream groups between Cholesky and triangular inverse, so the
    #      stages sum to slightly more than ms_per_step) ----
    stage_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(args.steps)]
    for evs in stage_ev:
        for e in evs:
            e.record()          # force creation of the underlying cudaEvent_t
    barrier()
    for s in range(args.steps):
        device_step(stage_ev[s])
    barrier()
    # [build, cholesky, trtri, solve (pre-contract), contract kernel, tail] in ms
    stages = np.array([[evs[i].elapsed_time(evs[i + 1]) for i in range(6)] for evs in stage_ev]).mean(axis=0)
    st = torch.tensor(stages, dtype=torch.float64, device=dev)

    # ---- e2e: public API, host parameters in, host objective + gradient out ----
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        f_e2e, g_e2e = model.loss_and_grad()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
    barrier()
    if world > 1:
        torch.distributed.all_reduce(t_dev, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(t_e2e, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(st, op=torch.distributed.ReduceOp.MAX)
    # ---- prediction (a7): latent means / variances at 1024 test points from the factor in the workspace
    n0 = 1024
    x0 = torch.as_tensor(np.random.default_rng(7).uniform(0, 1, (n0, d)))
    x0 = x0 * (model.x_max - model.x_min) + model.x_min
    model.predict(x0)                        # includes the aux refresh
    barrier()
    w0 = time.perf_counter()
    model.predict(x0)
    torch.cuda.synchronize()
    t_pred = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_pred, op=torch.distributed.ReduceOp.MAX)
    pred = {'n0': n0, 'wall_ms': float(t_pred) * 1e3, 'points_per_s': n0 / float(t_pred),
            'tflops': q * float(n0) * n * n / float(t_pred) / 1e12,
            'note': 'LCGP.predict() end to end (host x0 in, p x n0 host tensors out); flop = q n0 n^2'}
    clk = clocks.stop() if rank == 0 else None
    t_dev, t_e2e, stages = float(t_dev), float(t_e2e), st.cpu().numpy()

    if rank == 0:
        peaks, psrc = measured_peaks()
        fp64_peak, fp64_sustained = dgemm_peak(dev)
        npad = _cabi.padded(n)
        stage_flops = q_loc * n ** 3 / 3.0             # algorithmic flop of each of the three dense stages on one rank
        tf = lambda ms: stage_flops / (ms * 1e-3) / 1e12
        build_bytes = 8.0 * q_loc * n * (n + 1) / 2     # lower triangle written
        # the persistent kernel computes factor AND triangular inverse in one launch (then the trtri stage is empty)
        fused = stages[2] < 0.02 * stages[1]
        chol_flops = (2.0 if fused else 1.0) * stage_flops
        if fused and stages[1] > stages[4]:
            dom = dict(kernel='potrf_pll_kernel (persistent left-looking Cholesky + fused triangular inverse: TMA-staged DMMA tile '
                              'tasks with dependency flags, one launch)', ms=float(stages[1]), flops=chol_flops)
        else:
            dom = dict(kernel='gemm_tma_kernel<ContractJob> (A^-1 = U U^T tiles: TMA-staged DMMA GEMM + fused gradient contraction)',
                       ms=float(stages[4]), flops=stage_flops)
        dom_tf = dom['flops'] / (dom['ms'] * 1e-3) / 1e12
        traffic, traffic_source = None, None
        for tname in ('r2_pll_kernel_traffic.json', 'r2_contract_kernel_traffic.json', 'r1_contract_kernel_traffic.json'):
            tpath = os.path.join(ROOT, 'profiles', tname)
            if os.path.exists(tpath):                   # dram bytes of this launch from a COMMITTED ncu capture
                with open(tpath) as f:                  # (ncu cannot run inside the timed bench): not measured in this run
                    tj = json.load(f)
                if tj.get('q_loc') == q_loc and tj.get('n') == n and tj.get('kernel', 'ContractJob') in dom['kernel']:
                    traffic, traffic_source = tj['dram_bytes_per_launch'], f'ncu capture profiles/{tname}'
                    break
        line = {
            'metric': 'NLL+grad evals/s', 'value': args.steps / t_dev, 'unit': 'evals/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': t_dev / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_desc(args.config, n, d, p, q, model.submethod, world),
            'arm': 'lcgp_b200 CUDA path (liblcgp_b200.so through the C-ABI)',
            'e2e': {'value': args.steps / t_e2e, 'unit': 'evals/s', 'h2d_bytes_per_step': eng.h2d_bytes,
                    'd2h_bytes_per_step': eng.d2h_bytes if world == 1 else (1 + p + q * d + 2 * q + 1) * 8},
            'gpu_launches': launches,
            # dominant kernel = the largest single launch of the step, timed live with CUDA events recorded around it
            # inside the C-ABI call: the fused A^-1 / gradient-contraction GEMM (~1/3 of the step) when the Cholesky runs
            # as a launch chain (large batches), else the persistent Cholesky + inverse kernel (~2/3 of the step)
            'roofline': {'bound': 'tensor', 'kernel': dom['kernel'],
                         'achieved': dom_tf, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': dom_tf / fp64_peak,
                         'peak_source': 'cuBLAS DGEMM 8192^3 burst (best of 4) measured in this run; MEASURED_PEAKS.json has no '
                                        'FP64 figure and the profiling guide states no FP64 fallback',
                         'peak_sustained': fp64_sustained, 'frac_of_sustained': dom_tf / fp64_sustained,
                         'flops_per_launch': dom['flops'], 'kernel_ms': dom['ms'], 'traffic': traffic,
                         'traffic_source': traffic_source},
            'stages': {
                'build_ms': float(stages[0]), 'cholesky_ms': float(stages[1]), 'trtri_ms': float(stages[2]),
                'solve_ms': float(stages[3]), 'contract_kernel_ms': float(stages[4]), 'tail_ms': float(stages[5]),
                # build_A is NOT HBM bound: ncu (profiles/r2_ncu_full_pll_buildA.txt) shows the FP64 ALU pipe active 56 % of the
                # cycles and DRAM throughput at 14 % (~90 FP64 operations incl. an exp per 8-byte element); the GB/s figure
                # is kept because north_star asks for it
                'build_bound': 'fp64 ALU (ncu: sm__pipe_fp64_cycles_active 55.7 %, dram throughput 13.8 %); 1.3 % of the step',
                'build_gbs': build_bytes / (stages[0] * 1e-3) / 1e9, 'hbm_peak_gbs': peaks['hbm_gbs'],
                'hbm_peak_source': psrc, 'build_frac_of_hbm': build_bytes / (stages[0] * 1e-3) / 1e9 / peaks['hbm_gbs'],
                'cholesky_trtri_fused': bool(fused),       # True: cholesky_* below cover factor + inverse (2 n^3 / 3 flop per latent)
                'cholesky_tflops': chol_flops / (stages[1] * 1e-3) / 1e12,
                'cholesky_frac_of_dgemm': chol_flops / (stages[1] * 1e-3) / 1e12 / fp64_peak,
                'trtri_tflops': None if fused else tf(stages[2]),
                'trtri_frac_of_dgemm': None if fused else tf(stages[2]) / fp64_peak,
                'contract_tflops': tf(stages[4]),
                'whole_eval_tflops': q_loc * float(n) ** 3 / (t_dev / args.steps) / 1e12,
                'dgemm_peak_tflops': fp64_peak, 'dgemm_sustained_tflops': fp64_sustained, 'padded_n': npad},
            'clocks': clk, 'objective': f_e2e, 'grad_norm': float(np.linalg.norm(g_e2e)), 'ctor_s': t_ctor,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            try:
                _, _, dt = cpu_baseline_sample(x, y, mk, args.cpu_sample_latents, q, threads)
                per_eval = dt * q / args.cpu_sample_latents
                line['cpu_baseline'] = {'value': 1.0 / per_eval, 'unit': 'evals/s', 'cores': threads, 'kind': 'port',
                                        'sample': f'{args.cpu_sample_latents} of {q} latents at full n, forward + autograd '
                                                  f'backward, 1 repetition ({dt:.1f} s), extrapolated x{q // args.cpu_sample_latents}'}
            except Exception as ex:   # baseline is informational; never lose the GPU numbers
                line['cpu_baseline'] = {'value': None, 'unit': 'evals/s', 'cores': threads, 'kind': 'port',
                                        'sample': f'failed: {ex!r}'}
        else:
            line['cpu_baseline'] = None
        line['predict'] = pred
    # ---- fit wall time: every rank takes part when the fit is sharded ----
    fit_explicit = any(a.startswith('--fit-config') for a in sys.argv)
    if world > 1 and not fit_explicit:
        args.fit_config = 'workload'      # N > 1: the driver-timed fit is the sharded fit of the bench configuration itself
    fit_cfg = args.config if args.fit_config == 'workload' else args.fit_config
    if not args.no_fit:
        del model, eng
        torch.cuda.empty_cache()
        if rank == 0 and world == 1 and not args.no_named_configs:
            line['named_configs'] = named_configs(not args.no_cpu_baseline)
        cpu_pe = None
        if rank == 0 and fit_cfg == args.config and line.get('cpu_baseline') and line['cpu_baseline'].get('value'):
            cpu_pe = 1.0 / line['cpu_baseline']['value']
        fit = fit_wall(fit_cfg, not args.no_cpu_baseline and world == 1, cpu_pe, sharded=world > 1)
        if rank == 0:
            line['fit'] = fit
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def run_batched(args):
    """BASELINE config 5: 64 independent emulators (n=1024, d=6, p=64, q=8, seeds 1024..1087) fitted to convergence
    (SciPy L-BFGS-B defaults unless --maxiter), emulator i on rank i mod N ("replicas only": no collective during the
    fit); each rank advances its emulators in lock-step with one batched device call per step (lcgp_b200.batched).
    Prints its own JSON line (metric: emulator fits/s); not the headline bench line.  cpu_baseline (N = 1 only): the
    oracle timed on a few evaluations of ONE emulator on the host cores, extrapolated with the mean evaluation count."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    from lcgp_b200 import _cabi, fit_emulators, synthetic
    n_em = args.emulators
    data, mk = [], None
    for i in range(n_em):
        x, y, _, _, mk = synthetic.make_config('cfg5_one', seed=1024 + i)
        data.append((x, y))
    opts = dict(maxiter=args.maxiter) if args.maxiter > 0 else {}
    kw = dict(threads_per_gpu=args.threads, engine=args.engine)
    fit_emulators(data[:min(2 * world, n_em)], mk, fit_options=dict(maxiter=3), **kw)   # warm-up
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    l0 = int(_cabi.lib().lcgp_launch_count())
    t0 = time.perf_counter()
    res = fit_emulators(data, mk, fit_options=opts, **kw)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    dt = time.perf_counter() - t0
    launches = int(_cabi.lib().lcgp_launch_count()) - l0
    if rank == 0:
        evals = sum(r['n_evals'] for r in res)
        line = {'metric': 'emulator fits/s', 'value': n_em / dt, 'unit': 'fits/s', 'n_gpus': world,
                'evals_per_s': evals / dt, 'total_evals': evals, 'mean_evals_per_fit': evals / n_em, 'wall_s': dt,
                'higher_is_better': True, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': f'cfg5: {n_em} emulators n=1024 d=6 p=64 q=8 (full), L-BFGS-B ' +
                                       (f'maxiter={args.maxiter}' if args.maxiter > 0 else 'to convergence (SciPy defaults)'),
                           'engine': args.engine, 'emulators_per_rank': -(-n_em // world)},
                'scaling': 'replicas only', 'gpu_launches_rank0': launches,
                'converged': int(sum(bool(r.get('converged', True)) for r in res)),
                'mean_final_loss': float(np.mean([r['loss'] for r in res]))}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle.lcgp_oracle import LCGPOracle
                threads = os.cpu_count() or 1
                torch.set_num_threads(threads)
                o = LCGPOracle(y=data[0][1], x=data[0][0], skip_xnorm=True, **mk)
                o.loss_and_grad(o.neglpost_chol)
                t1 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    o.loss_and_grad(o.neglpost_chol)
                per_eval = (time.perf_counter() - t1) / reps
                line['cpu_baseline'] = {'value': 1.0 / (per_eval * evals / n_em), 'unit': 'fits/s', 'cores': threads, 'kind': 'port',
                                        'evals_per_s': 1.0 / per_eval,
                                        'sample': f'{reps} oracle evaluations (forward + autograd backward) of emulator 0 on '
                                                  f'{threads} host threads, {per_eval:.2f} s each; fits/s EXTRAPOLATED with the '
                                                  f'mean of {evals / n_em:.0f} evaluations per fit of the CUDA run'}
            except Exception as ex:
                line['cpu_baseline'] = {'value': None, 'sample': f'failed: {ex!r}'}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        try:
            run_reference(a)
        except Exception as ex:      # the driver expects one JSON line and exit code 0 from this arm
            if int(os.environ.get('RANK', '0')) == 0:
                print(json.dumps({'impl': 'reference', 'unavailable': f'oracle port failed on this host: {ex!r}'}))
    elif a.config == 'cfg5_batch':
        run_batched(a)
    else:
        run_ours(a)
