"""bench.py - CURIOUS training hot path on B200 (contract: see the repository prompt / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

metric   HER-relabelled transitions/s (BASELINE.json), Arm4-shaped synthetic replay:
         4 module buffers x 1e6 transitions (20 000 episodes x T=50, dimo=40, dimg=dimag=12, dimu=4, N=4).
step     ONE launch of the fused HER relabel+gather+reward+clip kernel over ROWS_PER_STEP rows
         (4096 batches of 256), LP-apportioned over the module buffers, Philox draws, producing the
         staged training batch (o, g, u, task_descr, o_2, r).
value    rows / device time (CUDA events), inputs resident in HBM, max over ranks, whole job.
e2e      the reference's training cycle through the public plugin API with HOST episode buffers:
         DDPG.store_episode(host episodes) + n_batches x DDPG.train() + DDPG.update_target_net()
         + device->host read of the critic losses; transitions consumed per second.
--impl reference   the oracle port of that same cycle on the host cores (the reference is pure Python;
         TF1/mpi4py/gym_flowers are not installable), one process per core up to 19.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_MODULES = 4
T = 50
BUFFER_TRANSITIONS = 1000000
BATCH = 256
ROWS_PER_STEP = 1 << 20
CP = [0.05, 0.2, 0.1, 0.0]
EPS_TASK = 0.4
N_BATCHES_E2E = 100            # config.py:72
REF_UPDATES_PER_STEP = 10      # bounded sample of the cycle for the CPU arm


def algorithmic_bytes_per_transition(dims, n_modules, g_len=3):
    """reads o,o_2,u,td,g and the module slices of ag_2 / future ag; writes o,o_2,g,u,td,r."""
    rd = 4 * (2 * dims['o'] + dims['u'] + n_modules + dims['g'] + 2 * g_len)
    wr = 4 * (2 * dims['o'] + dims['g'] + dims['u'] + n_modules + 1)
    return rd + wr


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md's clocks line).  Sampled through NVML
    in-process every 4 ms (the HER region lasts ~10 ms; one `nvidia-smi` call takes longer than that), falling back to
    an `nvidia-smi --query-gpu` loop when pynvml is unavailable."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    REASON_BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20),
                   ('sw_power_cap', 0x4))

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.source = 'nvidia-smi'
        self._nvml = self._handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            handle = None
            try:                                     # CUDA_VISIBLE_DEVICES may renumber: go by UUID when torch exposes it
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self._nvml, self._handle, self.source = pynvml, handle, 'nvml'
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        return [str(sm), str(mx)] + ['Active' if mask & bit else 'Not Active' for _, bit in self.REASON_BITS]

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.check_output(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                                   '--format=csv,noheader,nounits'], timeout=5).decode().strip()
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self._stop_evt.wait(0.004 if self._nvml is not None else 0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[2:]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': self.source}


def her_traffic_per_launch(rows_per_step):
    """DRAM bytes of one fused HER launch from the committed `ncu --set full` capture (profiles/), scaled to the
    rows of this launch; None if the capture is missing."""
    p = os.path.join(ROOT, 'profiles', 'r01_her_traffic.json')
    try:
        t = json.load(open(p))
        return (t['dram_bytes_read'] + t['dram_bytes_write']) * rows_per_step / t['rows_per_launch']
    except Exception:
        return None


def measured_peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------
# workload construction
# ------------------------------------------------------------------------------------------------
def synth_dims():
    from curious_b200 import synth
    dims = synth.arm_dims(N_MODULES)
    ag_ids, g_ids = synth.arm_task_ids(N_MODULES)
    return dims, ag_ids, g_ids


def fill_buffer_on_device(buf, dims, seed):
    """Fill a ReplayBuffer to capacity with Arm-shaped synthetic episodes generated ON the device (same value
    structure as curious_b200.synth.make_episodes; input generation only)."""
    import torch
    from curious_b200.replay_buffer import StagedEpisodes
    g = torch.Generator(device=buf.device)
    g.manual_seed(seed)
    E, N = buf.size, dims['task_descr']
    chunk = 2000
    dev = buf.device
    for e0 in range(0, E, chunk):
        n = min(chunk, E - e0)
        o = torch.randn((n, T + 1, dims['o']), generator=g, device=dev).clamp_(-5, 5)
        ag0 = (torch.rand((n, 1, dims['ag']), generator=g, device=dev) - 0.5) * 0.3
        moving = (torch.rand((n, N), generator=g, device=dev) >= 0.3).float().repeat_interleave(dims['ag'] // N, dim=1)
        steps = torch.randn((n, T, dims['ag']), generator=g, device=dev) * 0.02 * moving[:, None, :]
        ag = torch.cat([ag0, ag0 + torch.cumsum(steps, dim=1)], dim=1).contiguous()
        task = torch.randint(0, N, (n,), generator=g, device=dev)
        td = torch.nn.functional.one_hot(task, N).float()[:, None, :].expand(n, T, N).contiguous()
        gv = (torch.rand((n, dims['g']), generator=g, device=dev) - 0.5) * 0.3
        gmask = td[:, 0, :].repeat_interleave(dims['g'] // N, dim=1)
        gg = (gv * gmask)[:, None, :].expand(n, T, dims['g']).contiguous()
        u = torch.rand((n, T, dims['u']), generator=g, device=dev) * 2 - 1
        change = ((ag[:, :1, :] - ag[:, 1:, :]).abs() > 1e-3).float().contiguous()
        info = (torch.rand((n, T, 1), generator=g, device=dev) < 0.25).float()
        staged = StagedEpisodes.from_device(dict(o=o.contiguous(), ag=ag, g=gg, u=u.contiguous(), task_descr=td,
                                                 change=change, info=info), buf.layout)
        staged.store([(i, buf.storage, buf.cold, e0 + i) for i in range(n)])
        torch.cuda.synchronize()
    buf.current_size = E
    buf.n_transitions_stored = E * T


def build_gpu_workload(device, seed):
    from curious_b200 import her, synth
    from curious_b200.ddpg import DDPG
    from curious_b200.replay_buffer import ReplayBuffer
    from curious_b200.reward import ModuleDistanceReward
    dims, ag_ids, g_ids = synth_dims()
    sampler = her.make_sample_multi_task_her_transitions('her', 4, 'replay_task_cp_buffer',
                                                         ModuleDistanceReward(ag_ids, g_ids), tasks_ag_id=ag_ids,
                                                         tasks_g_id=g_ids)
    sampler.rng = 'philox'
    sampler.seed = seed
    shapes = synth.buffer_shapes(dims, T)
    buffers = [ReplayBuffer(shapes, BUFFER_TRANSITIONS if i > 0 else T, T, sampler, device=device)
               for i in range(N_MODULES + 1)]          # buffer 0 is never written by the reference (ddpg.py:191)
    for i in range(1, N_MODULES + 1):
        fill_buffer_on_device(buffers[i], dims, seed * 100 + i)
    gamma = 1. - 1. / T

    def make_agent(batch_size=BATCH, structure='curious', task_replay='replay_task_cp_buffer', **extra):
        a = DDPG(input_dims=dims, hidden=256, layers=3, network_class='baselines.her.actor_critic:MultiTaskActorCritic',
                 polyak=0.95, batch_size=batch_size, Q_lr=0.001, pi_lr=0.001, norm_eps=0.01, norm_clip=5, max_u=1.,
                 action_l2=1.0, clip_obs=200., scope='ddpg', T=T, rollout_batch_size=2,
                 subtract_goals=lambda a, b: a - b, relative_goals=False, clip_pos_returns=True,
                 clip_return=1. / (1. - gamma), normalize_obs=False, sample_transitions=sampler, gamma=gamma,
                 buffers=buffers, tasks_ag_id=ag_ids, tasks_g_id=g_ids, task_replay=task_replay,
                 eps_task=EPS_TASK, structure=structure, her_rng='philox', seed=0, device=device,
                 grad_exchange=os.environ.get('CUR_GRAD_EXCHANGE', 'auto'), **extra)
        a.cp = np.array(CP)
        return a

    agent = make_agent()
    agent.make_agent = make_agent
    return agent, sampler, buffers, dims, ag_ids, g_ids


def update_flops(dims, n_modules, batch, hidden=256):
    """algorithmic FLOPs of one DDPG update (forward of main/target pi and the three Q passes, critic,
    actor-through-critic and actor backward chains)."""
    H, B = hidden, batch
    s_pi, s_q = dims['o'] + n_modules, dims['o'] + n_modules + dims['u']
    f_pi = 2 * B * ((s_pi + dims['g']) * H + 2 * H * H + H * dims['u'])
    f_q = 2 * B * ((s_q + dims['g']) * H + 2 * H * H + H)
    fwd = 2 * f_pi + 3 * f_q
    bwd = (2 * f_q - 2 * B * (s_q + dims['g']) * H) + (f_q - 2 * B * (s_q + dims['g'] - dims['u']) * H) + \
          (2 * f_pi - 2 * B * (s_pi + dims['g']) * H)
    return fwd + bwd


def time_updates(fn, n, torch):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def large_batch_sweep(agent, dims, torch):
    """BASELINE config 5: per-GPU batch sweep of the DDPG update through DDPG.train() (HER sample + grads + Adam in
    one CUDA graph) and structure='task_experts' as grouped launches.  At batch >= 1024 the hidden-layer GEMMs run on
    tcgen05 (3xTF32): `tensor_frac` = 3 x algorithmic FLOP/s / (measured dense bf16 peak / 2), i.e. against the TF32
    rate of the tensor pipe, all 3 error-compensation passes counted."""
    from curious_b200 import _lib
    from curious_b200.experts import TaskExperts
    import ctypes as C
    try:
        bf16 = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops'])
    except Exception:
        bf16 = 2250.0 * 0.745
    sweep = []
    for b in (256, 512, 1024, 2048, 4096, 16384):
        a = agent if b == BATCH else agent.make_agent(batch_size=b)
        ms = time_updates(a.train, 200 if b <= 1024 else 40, torch)
        fl = update_flops(dims, N_MODULES, b)
        tc = bool(_lib.load().cur_ddpg_uses_tensor_cores(C.byref(a.net.desc), b)) and not a._use_rows(b)
        sweep.append({'batch': b, 'update_us': 1e3 * ms, 'updates_per_s': 1e3 / ms, 'transitions_per_s': 1e3 * b / ms,
                      'tflops': fl / (ms * 1e-3) / 1e12, 'schedule': 'rows' if a._use_rows(b) else 'levels',
                      'tensor_cores': tc, 'tensor_frac': (3 * fl / (ms * 1e-3) / 1e12) / (bf16 / 2) if tc else None})
        if b != BATCH:
            del a
            torch.cuda.empty_cache()
    experts = []
    for b in (256, 4096):
        ps = [agent.make_agent(batch_size=b, structure='task_experts', task_replay='replay_current_task_buffer', t_id=t,
                               update_schedule='auto') for t in range(N_MODULES)]
        for p in ps:
            p.cp = np.array(CP)
        grp = TaskExperts(ps)
        ms = time_updates(grp.train, 100 if b <= 1024 else 20, torch)
        experts.append({'experts': N_MODULES, 'batch': b, 'round_us': 1e3 * ms, 'expert_updates_per_s': 1e3 * N_MODULES / ms,
                        'mode': 'sequential rows' if all(p._use_rows(b) for p in ps) else 'grouped levels'})
        del grp, ps
        torch.cuda.empty_cache()
    return {'batch_sweep': sweep, 'task_experts': experts, 'tf32_peak_tflops_assumed': bf16 / 2}


def her_arm8_line(device, torch):
    """BASELINE config 4 shape: MultiTaskFetchArm8-v5 (4 distractor modules; buffers 6..8 alias 5, ddpg.py:107-110),
    dimo=64, dimg=dimag=24, N=8 -> 1340 algorithmic bytes per transition.  Same fused kernel, same launch."""
    from curious_b200 import apportion, her, synth
    from curious_b200.replay_buffer import ReplayBuffer
    from curious_b200.reward import ModuleDistanceReward
    n_mod = 8
    dims = synth.arm_dims(n_mod)
    ag_ids, g_ids = synth.arm_task_ids(n_mod)
    sampler = her.make_sample_multi_task_her_transitions('her', 4, 'replay_task_cp_buffer',
                                                         ModuleDistanceReward(ag_ids, g_ids), tasks_ag_id=ag_ids,
                                                         tasks_g_id=g_ids)
    sampler.rng = 'philox'
    sampler.seed = 5
    shapes = synth.buffer_shapes(dims, T)
    buffers = [ReplayBuffer(shapes, BUFFER_TRANSITIONS if 0 < i <= 5 else T, T, sampler, device=device) for i in range(6)]
    for i in range(1, 6):
        fill_buffer_on_device(buffers[i], dims, 700 + i)
    buffers += [buffers[5]] * 3                                       # distractor modules share one buffer
    cp = np.array([0.05, 0.2, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0])
    sizes = [b.current_size for b in buffers]
    prop = apportion.proportions_curious(sizes, T, ROWS_PER_STEP, 'replay_task_cp_buffer', cp, EPS_TASK)
    segs = [(buffers[i].device_view(), int(prop[i]), i - 1) for i in range(1, len(buffers)) if prop[i] > 0]
    out = {}
    want = ('o', 'g', 'u', 'td', 'o_2', 'r')
    ms = time_updates(lambda: sampler.sample_device(segs, ROWS_PER_STEP, clip_obs=200.0, want=want, out=out), 20, torch)
    bpt = algorithmic_bytes_per_transition(dims, n_mod)
    peak, _ = measured_peak_hbm()
    achieved = bpt * ROWS_PER_STEP / (ms * 1e-3) / 1e9
    del buffers, out
    torch.cuda.empty_cache()
    return {'transitions_per_s': ROWS_PER_STEP / (ms * 1e-3), 'ms_per_launch': ms, 'algorithmic_bytes_per_transition': bpt,
            'achieved_gbs': achieved, 'frac_of_hbm_peak': achieved / peak,
            'workload': 'arm8-shaped: 5 distinct module buffers x 1e6 transitions (dimo=64, dimg=24, dimu=4, N=8), '
                        '%d rows per launch' % ROWS_PER_STEP}


def her_step_segments(buffers, rows):
    from curious_b200 import apportion
    sizes = [b.current_size for b in buffers]
    prop = apportion.proportions_curious(sizes, T, rows, 'replay_task_cp_buffer', np.array(CP), EPS_TASK)
    return [(buffers[i].device_view(), int(prop[i]), i - 1) for i in range(1, len(buffers)) if prop[i] > 0]


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) - used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------------
def build_cpu_workload(seed, episodes_per_buffer):
    from curious_b200 import synth
    from oracle import ddpg_oracle, her_oracle, replay_oracle
    from oracle.reward_oracle import ModuleDistanceReward
    dims, ag_ids, g_ids = synth_dims()
    sampler = her_oracle.make_sample_multi_task_her_transitions('her', 4, 'replay_task_cp_buffer',
                                                                ModuleDistanceReward(ag_ids, g_ids),
                                                                tasks_ag_id=ag_ids, tasks_g_id=g_ids)
    shapes = synth.buffer_shapes(dims, T)
    buffers = [replay_oracle.ReplayBufferOracle(shapes, (episodes_per_buffer if i > 0 else 1) * T, T, sampler)
               for i in range(N_MODULES + 1)]
    rng = np.random.RandomState(seed)
    for i in range(1, N_MODULES + 1):
        for e0 in range(0, episodes_per_buffer, 2000):
            n = min(2000, episodes_per_buffer - e0)
            buffers[i].store_episode(synth.make_episodes(rng, n, T, dims))
    gamma = 1. - 1. / T
    agent = ddpg_oracle.DDPGOracle(
        input_dims=dims, hidden=256, layers=3, polyak=0.95, batch_size=BATCH, Q_lr=0.001, pi_lr=0.001, norm_eps=0.01,
        norm_clip=5, max_u=1., action_l2=1.0, clip_obs=200., T=T, rollout_batch_size=2, relative_goals=False,
        clip_pos_returns=True, clip_return=1. / (1. - gamma), normalize_obs=False, sample_transitions=sampler,
        gamma=gamma, buffers=buffers, structure='curious', tasks_ag_id=ag_ids, tasks_g_id=g_ids,
        task_replay='replay_task_cp_buffer', eps_task=EPS_TASK, weights_rng=np.random.RandomState(0))
    agent.cp = np.array(CP)
    return agent, dims


def cpu_sampler_baseline(seconds=12.0):
    """The oracle's DDPG.sample_batch (apportion + per-buffer HER sampler + concat + shuffle + preprocess) on one
    core, full-size buffers: transitions/s.  kind = 'port' (restatement of her.py / replay_buffer.py / ddpg.py)."""
    import contextlib
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except Exception:
        ctx = contextlib.nullcontext()
    with ctx:
        agent, _ = build_cpu_workload(0, BUFFER_TRANSITIONS // T)
        np.random.seed(0)
        for _ in range(20):
            agent.sample_batch()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            agent.sample_batch()
            n += 1
        dt = time.perf_counter() - t0
    return {'value': n * BATCH / dt, 'unit': 'transitions/s', 'cores': 1, 'kind': 'port',
            'sample': '%d sample_batch calls x %d rows on 4 x 1e6-transition float64 buffers, 1 thread '
                      '(oracle port of her.py/replay_buffer.py/ddpg.py sample_batch)' % (n, BATCH)}


def _ref_worker(rank, n_steps, n_warm, updates, episodes_per_buffer, q):
    os.environ['OMP_NUM_THREADS'] = '1'
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    agent, dims = build_cpu_workload(rank, episodes_per_buffer)
    from curious_b200 import synth
    rng = np.random.RandomState(1000 + rank)
    np.random.seed(1000000 * rank)
    q.put(('ready', rank))
    times = []
    for s in range(n_warm + n_steps):
        ep = synth.make_episodes(rng, 2, T, dims, change_dtype=bool)
        t0 = time.perf_counter()
        agent.store_episode(ep, np.array(CP), 2 * (s + 1))
        for _ in range(updates):
            agent.train()
        agent.update_target_net()
        times.append(time.perf_counter() - t0)
    q.put(('done', rank, times[n_warm:]))


def run_reference_arm(args):
    """The training cycle of the reference restated on the CPU (oracle port), one single-threaded process per
    worker like the reference's `mpirun -np 19 --bind-to core` (train.py:221-231); no all-reduce is modelled, so
    this over-estimates the CPU arm.  Each step is a bounded sample of the cycle: REF_UPDATES_PER_STEP updates."""
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(19, cores))
    episodes = BUFFER_TRANSITIONS // T
    # bound memory: each worker owns 4 float64 buffers of 0.69 GB
    try:
        import psutil
        avail = psutil.virtual_memory().available
        procs = max(1, min(procs, int(avail * 0.6 // (4 * 0.7e9))))
    except Exception:
        pass
    ctx = mp.get_context('fork')
    q = ctx.Queue()
    workers = [ctx.Process(target=_ref_worker, args=(r, args.steps, args.warmup, REF_UPDATES_PER_STEP, episodes, q))
               for r in range(procs)]
    for w in workers:
        w.start()
    done = {}
    while len(done) < procs:
        msg = q.get()
        if msg[0] == 'done':
            done[msg[1]] = msg[2]
    for w in workers:
        w.join()
    per_step = np.max(np.array([done[r] for r in range(procs)]), axis=0)      # slowest worker per step
    total = float(per_step.sum())
    value = procs * REF_UPDATES_PER_STEP * BATCH * args.steps / total
    dims, _, _ = synth_dims()
    line = {
        'impl': 'reference', 'metric': 'HER-relabelled transitions/s', 'value': value, 'unit': 'transitions/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'arm4-shaped: 4 module buffers x 1e6 transitions (T=50, dimo=40, dimg=12, dimu=4, N=4), '
                               'batch 256, replay_task_cp_buffer; CPU arm = training cycle store_episode + %d x train + '
                               'update_target_net per step' % REF_UPDATES_PER_STEP},
        'cpu_baseline': {'value': value, 'unit': 'transitions/s', 'cores': procs, 'kind': 'port',
                         'sample': '%d single-threaded worker processes x %d steps x %d updates of batch %d (oracle '
                                   'NumPy port of her.py/replay_buffer.py/ddpg.py/mpi_adam.py; TF1 graph restated in '
                                   'float32 NumPy/BLAS)' % (procs, args.steps, REF_UPDATES_PER_STEP, BATCH)},
        'e2e': {'value': value, 'unit': 'transitions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'updates_per_s': procs * REF_UPDATES_PER_STEP * args.steps / total,
        'host_cores': cores,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    from curious_b200 import _lib
    _lib.load()
    agent, sampler, buffers, dims, ag_ids, g_ids = build_gpu_workload(device, seed=1 + rank)
    segs = her_step_segments(buffers, ROWS_PER_STEP)
    want = ('o', 'g', 'u', 'td', 'o_2', 'r')
    out = {}

    def her_step():
        return sampler.sample_device(segs, ROWS_PER_STEP, clip_obs=200.0, want=want, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        her_step()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        her_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    # ---- DDPG updates/s: device-timed train() (sample + grads + all-reduce + Adam) through the public API
    for _ in range(5):
        agent.train()
    barrier()
    n_upd = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_upd):
        agent.train()
    e1.record()
    barrier()
    upd_ms = e0.elapsed_time(e1)
    # ---- 19-worker-equivalent (BASELINE config 4 / SURVEY 8e): the reference's 19 MPI workers spread over the ranks,
    # ceil(19 / world) batch-256 workers per GPU, gradients summed per rank and across ranks, one Adam step
    k19 = -(-19 // world)
    agent19 = agent.make_agent(workers_per_rank=k19)
    for _ in range(3):
        agent19.train()
    barrier()
    n19 = 30
    e0.record()
    for _ in range(n19):
        agent19.train()
    e1.record()
    barrier()
    upd19_ms = e0.elapsed_time(e1)
    del agent19
    # ---- e2e: training cycles through the plugin API with host episode buffers
    from curious_b200 import synth
    rng = np.random.RandomState(99 + rank)
    host_eps = [synth.make_episodes(rng, 2, T, dims, change_dtype=bool) for _ in range(8)]
    h2d = sum(np.asarray(v).size * 4 for v in host_eps[0].values())

    def cycle(i):
        agent.store_episode({k: v for k, v in host_eps[i % len(host_eps)].items()}, np.array(CP), 2 * (i + 1))
        losses = [agent.train()[0] for _ in range(N_BATCHES_E2E)]
        agent.update_target_net()
        return torch.stack([l.tensor for l in losses]).cpu().numpy()          # device->host read of the results

    for i in range(3):
        cycle(i)
    barrier()
    n_cyc = 10
    t0 = time.perf_counter()
    for i in range(n_cyc):
        res = cycle(3 + i)
    barrier()
    cyc_s = time.perf_counter() - t0
    clock_info = clocks.stop()
    if world > 1:
        tt = torch.tensor([ms, upd_ms, cyc_s, upd19_ms], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, upd_ms, cyc_s, upd19_ms = [float(x) for x in tt.cpu()]
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bpt = algorithmic_bytes_per_transition(dims, N_MODULES)
        kernel_ms = ms / args.steps                       # one step == one launch of the fused kernel
        achieved = bpt * ROWS_PER_STEP / (kernel_ms * 1e-3) / 1e9
        line = {
            'metric': 'HER-relabelled transitions/s', 'value': world * ROWS_PER_STEP * args.steps / (ms * 1e-3),
            'unit': 'transitions/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': kernel_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'arm4-shaped: 4 module buffers x 1e6 transitions (T=50, dimo=40, dimg=12, dimu=4, '
                                   'N=4), LP-apportioned (replay_task_cp_buffer, cp=%s), %d rows per fused launch '
                                   '(= %d batches of 256), Philox draws' % (CP, ROWS_PER_STEP, ROWS_PER_STEP // BATCH),
                       'l2': 'inputs (1.44 GB of replay rows per rank) and outputs (0.42 GB) exceed the 126 MB L2',
                       'rows_per_step': ROWS_PER_STEP, 'batch': BATCH},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': her_traffic_per_launch(ROWS_PER_STEP), 'peak_source': peak_src,
                         'algorithmic_bytes': bpt * ROWS_PER_STEP, 'algorithmic_bytes_per_transition': bpt,
                         'kernel': 'her_sample_kernel'},
            'e2e': {'value': world * n_cyc * N_BATCHES_E2E * BATCH / cyc_s, 'unit': 'transitions/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(res.nbytes),
                    'what': 'DDPG.store_episode(host) + %d x DDPG.train() + update_target_net + loss readback per cycle'
                            % N_BATCHES_E2E, 'cycle_ms': 1e3 * cyc_s / n_cyc,
                    'updates_per_s': world * n_cyc * N_BATCHES_E2E / cyc_s},
            'gpu_launches': args.steps,
            'clocks': clock_info,
            'updates_per_s': world * n_upd / (upd_ms * 1e-3),
            'update_us': 1e3 * upd_ms / n_upd,
            'update_schedule': 'rows' if agent._use_rows(BATCH) else 'levels',
            'workers19_equivalent': {'workers_per_rank': k19, 'workers': k19 * world, 'global_batch': k19 * world * BATCH,
                                     'update_us': 1e3 * upd19_ms / n19, 'updates_per_s': n19 / (upd19_ms * 1e-3),
                                     'transitions_per_s': k19 * world * BATCH * n19 / (upd19_ms * 1e-3)},
        }
        if world == 1 and not args.no_sweep:
            line['ddpg_update'] = large_batch_sweep(agent, dims, torch)
            line['her_arm8'] = her_arm8_line(device, torch)
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_sampler_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sweep', action='store_true', help='skip the batch sweep / task_experts section')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
