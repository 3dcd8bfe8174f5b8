"""structure='task_experts' as one grouped update - the B200 form of the reference's per-module experts.

The reference builds one DDPG per module over shared per-module buffers (train.py:287-289, ddpg.py:302-318) and
trains ONE of them per epoch (`policy[i_policy].train()`, train.py:108-113).  On a B200 a batch-256 update of one
expert leaves the GPU mostly idle (every GEMM level is latency-bound), so `TaskExperts.train()` steps ALL experts
in one sequence of grouped launches: every dependency level of the DDPG graph is one launch over the problems of all
experts (csrc/ddpg.cu cur_ddpg_grads_group; FFMA grouped GEMM at batch 256, the tcgen05 batch at batch >= 1024).
Semantically it is `for p in policies: p.train()` - every expert keeps its own parameters, Adam state, Philox
stream, LP apportioning and buffers, and the result is bit-identical to stepping them one after the other on the
levels schedule.
"""
import ctypes as C

import torch

from . import _lib
from .parallel import allreduce_sum_, world as _world
from .util import LazyHost, capture_graph


class TaskExperts(object):
    def __init__(self, policies, use_cuda_graph=True, mode='auto'):
        """mode: 'grouped' = grouped launches; 'sequential' = p.train() one after the other; 'auto' = sequential
        when every expert runs the rows schedule at batch <= 512 (one expert's row CTAs already fill the 148 SMs;
        measured per round of 4 experts: batch 256 237 us vs 322 us grouped, 512 435 vs 524, 1024 841 vs 372),
        grouped otherwise."""
        assert len(policies) >= 1 and mode in ('auto', 'grouped', 'sequential')
        self.mode = mode
        p0 = policies[0]
        for p in policies:
            assert bytes(p.net.desc) == bytes(p0.net.desc), 'experts must share the network shape'
            assert p.batch_size == p0.batch_size and p.device == p0.device
        self.policies = list(policies)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._experts = None

    def __len__(self):
        return len(self.policies)

    def __getitem__(self, i):
        return self.policies[i]

    # ------------------------------------------------------------------------------------------
    def _expert_array(self, graph):
        arr = (_lib.DdpgExpert * len(self.policies))()
        for e, p in zip(arr, self.policies):
            n = p.batch_size
            b = p._gbatch if graph else p._batch
            g2 = b['g_2'] if (not graph or p.relative_goals) else b['g']
            e.theta_main = p.theta_main.data_ptr()
            e.theta_target = p.theta_target.data_ptr()
            e.stats = p._stats
            e.has_stats = 1
            e.batch = _lib.Batch(b['o'].data_ptr(), b['g'].data_ptr(), b['u'].data_ptr(),
                                 b['td'].data_ptr() if p.modular else None, b['o_2'].data_ptr(), g2.data_ptr(),
                                 b['r'].data_ptr(), n)
            e.hyper = p._ghyper if graph else p._hyper
            e.workspace = p._workspace(n).data_ptr()
            e.grads = p.grads.data_ptr()
            if graph:
                e.q_loss, e.pi_loss, e.q_pi = p._q_ring.data_ptr(), p._pi_ring.data_ptr(), p._q_pi.data_ptr()
            else:
                p._q_pi_eager = torch.empty((n, 1), dtype=torch.float32, device=p.device)
                e.q_loss, e.pi_loss, e.q_pi = p._q_loss.data_ptr(), p._pi_loss.data_ptr(), p._q_pi_eager.data_ptr()
        return arr

    def _launch(self, graph):
        """HER sample per expert, grouped gradients, then (single rank) Adam per expert."""
        lib = _lib.load()
        for p in self.policies:
            if graph:
                segs = [(buf.device_view(), 0, ttr) for buf, ttr in p._all_segments()]
                p.sample_transitions.sample_device(segs, p.batch_size, clip_obs=p.clip_obs,
                                                   relative_goals=p.relative_goals, want=p._gwant, out=p._gbatch,
                                                   dyn=p._dyn_dev.data_ptr(), call_offset=p.GRAPH_STREAM_OFFSET)
            else:
                p.stage_batch()
        arr = self._expert_array(graph)
        self._keep = arr
        _lib.check(lib.cur_ddpg_grads_group(_lib.stream_ptr(), C.byref(self.policies[0].net.desc), len(arr), arr),
                   'cur_ddpg_grads_group')

    def _build_graph(self):
        for p in self.policies:
            p.grad_exchange = 'nccl'          # several ranks: NCCL all-reduce + Adam after the graph
            p._prepare_graph_state()
            p._workspace(p.batch_size)
        dev = self.policies[0].device
        torch.cuda.current_stream().synchronize()
        states = [(p.theta_main.clone(), p._adam_m.clone(), p._adam_v.clone(), p._step.clone()) for p in self.policies]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up outside capture
            self._launch(True)
            for p in self.policies:
                p._launch_adam()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        for p, st in zip(self.policies, states):
            p.theta_main.copy_(st[0])
            for dst, src in zip((p._adam_m, p._adam_v, p._step), st[1:]):
                dst.copy_(src)
        self._single_rank = _world(self.policies[0].comm)[1] == 1
        g = torch.cuda.CUDAGraph()
        with capture_graph(g):
            self._launch(True)
            if self._single_rank:
                for p in self.policies:
                    p._launch_adam()
        self._graph = g

    def train(self, stage=True):
        """One update of every expert.  Returns [(critic_loss, actor_loss), ...] like DDPG.train (ddpg.py:368-373)."""
        ps = self.policies
        if self.mode == 'sequential' or (self.mode == 'auto' and all(
                p.update_schedule != 'levels' and p.batch_size <= 512 and p._use_rows(p.batch_size) for p in ps)):
            return [p.train(stage) for p in ps]
        graph = stage and self.use_cuda_graph and all(p.her_rng == 'philox' for p in ps)
        if graph:
            if self._graph is None:
                self._build_graph()
            for p in ps:
                if p.Q_adam.t % 100 == 0:
                    p.Q_adam.check_synced()
                    p.pi_adam.check_synced()
                p._refresh_dyn()
            slots = [p._n_updates % p.LOSS_RING for p in ps]
            self._graph.replay()
            if not self._single_rank:
                for p in ps:
                    allreduce_sum_(p.grads, p.comm)          # SUM, not mean (ddpg.py:452-453)
                    p._launch_adam()
            out = []
            for p, slot in zip(ps, slots):
                p._n_updates += 1
                p.Q_adam.t += 1
                p.pi_adam.t += 1
                out.append((LazyHost(p._q_ring[slot]), LazyHost(p._q_pi)))
            return out
        if stage:
            self._launch(False)
        else:
            arr = self._expert_array(False)
            self._keep = arr
            _lib.check(_lib.load().cur_ddpg_grads_group(_lib.stream_ptr(), C.byref(ps[0].net.desc), len(arr), arr),
                       'cur_ddpg_grads_group')
        out = []
        for p in ps:
            p._update(p._view(p.grads, 'Q'), p._view(p.grads, 'pi'))
            p._step += 1
            p._n_updates += 1
            out.append((LazyHost(p._q_loss.clone().reshape(())), LazyHost(p._q_pi_eager)))
        return out

    def update_target_net(self):
        for p in self.policies:
            p.update_target_net()
