"""RolloutWorker - the caller of DDPG.get_actions / the producer of DDPG.store_episode's input.

Mirror of reference baselines/her/rollout.py:13-491 (same constructor arguments, same methods, same episode
dict) for environments with the gym_flowers attribute contract (curious_b200.envs.ModularPointEnv here, since
gym_flowers / MuJoCo are absent).  What changed:
  * one batched `policy.get_actions` device forward per timestep for all `rollout_batch_size` environments
    (rollout.py:209-226 already batches per worker; with many envs per GPU rank this is north_star's "batched
    get_actions on device, MuJoCo stepping on host cores"),
  * the rank-0 LP block (rollout.py:316-404) lives in curious_b200.queues.CompetenceTracker; the MPI gathers /
    scatters / broadcasts become torch.distributed object collectives,
  * SAGG-RIAC goal selection (`goal_selection='active'`) is not supported (readme.md:19 marks it unsupported).
"""
import pickle
from collections import deque

import numpy as np

from .parallel import rank as _rank, world as _world
from .queues import CompetenceTracker
from .util import convert_episode_to_batch_major, store_args


class RolloutWorker(object):
    @store_args
    def __init__(self, make_env, policy, dims, logger, T, rollout_batch_size=1, exploit=False, use_target_net=False,
                 compute_Q=False, noise_eps=0, random_eps=0, history_len=100, render=False, structure='curious',
                 task_selection='random', goal_selection='random', queue_length=500, eval=False, unique_task=None,
                 temperature=None, **kwargs):
        """See reference rollout.py:19-41 for the arguments."""
        assert goal_selection != 'active', "goal_selection='active' (SAGG-RIAC) is unsupported, as in the reference readme"
        self.comm = kwargs.get('comm')
        self.envs = [make_env() for _ in range(rollout_batch_size)]
        assert self.T > 0
        self.info_keys = [key.replace('info_', '') for key in dims.keys() if key.startswith('info_')]
        self.success_history = deque(maxlen=history_len)
        self.reward_history = deque(maxlen=history_len)
        self.Q_history = deque(maxlen=history_len)
        self.n_episodes = 0
        self.g = np.empty((self.rollout_batch_size, self.dims['g']), np.float32)
        self.initial_o = np.empty((self.rollout_batch_size, self.dims['o']), np.float32)
        self.initial_ag = np.empty((self.rollout_batch_size, self.dims['ag']), np.float32)
        self.rank = _rank(self.comm)
        self.nb_cpu = _world(self.comm)[1]
        self.nb_goals_per_rollout = self.nb_cpu * self.rollout_batch_size
        self.nb_tasks = self.envs[0].unwrapped.nb_tasks
        self.C = np.zeros([self.nb_tasks])
        self.CP = np.zeros([self.nb_tasks])
        self.modular = self.structure in ('curious', 'task_experts')
        if self.modular:
            self.tasks_ag_id = self.envs[0].unwrapped.tasks_ag_id
            self.tasks_g_id = self.envs[0].unwrapped.tasks_g_id
            self.task_descr = np.empty((self.rollout_batch_size, self.nb_tasks), np.float32)
            self.tracker = CompetenceTracker(self.nb_tasks, queue_length=queue_length, task_selection=task_selection,
                                             structure=structure, unique_task=unique_task, eval=eval, comm=self.comm)
            self.p = self.tracker.p.copy()
            self.competence_computers = self.tracker.competence_computers
            self.task_history = deque()
            self.goal_history = deque()
        elif self.structure == 'flat':
            for env in self.envs:
                env.unwrapped.set_flat_env()
        self.stochastic_reset = False
        self.count = -1
        self.reset_all_rollouts()
        self.clear_history()

    # ------------------------------------------------------------------------------------------------------
    def _scatter(self, per_rank_values):
        """rank 0's list (one entry per rank) -> this rank's entry (MPI.COMM_WORLD.scatter, rollout.py:139-140)."""
        if self.nb_cpu == 1:
            return per_rank_values[0]
        import torch.distributed as dist
        box = [per_rank_values if self.rank == 0 else None]
        group = _world(self.comm)[0]
        dist.broadcast_object_list(box, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
        return box[0][self.rank]

    def reset_rollout(self, i):
        """rollout.py:104-165: reset env i, sample the next task (by p) and goal on rank 0, hand them to the env."""
        env = self.envs[i].unwrapped
        if self.eval or not self.stochastic_reset or np.random.rand() < 0.3 or self.exploit:
            self.envs[i].reset()
        if self.modular:
            tasks, goals = [], []
            if self.rank == 0:
                tasks = np.random.choice(range(self.nb_tasks), p=self.p, size=self.nb_cpu).tolist()
                goals = [np.random.uniform(-1, 1, len(self.tasks_g_id[tasks[c]])) for c in range(self.nb_cpu)]
                for cpu in range(self.nb_cpu):
                    good_ind = cpu * self.rollout_batch_size + i
                    self.tasks[good_ind] = tasks[cpu]
                    self.goals[good_ind] = env._compute_goal(goals[cpu], tasks[cpu], eval=self.eval)[0][
                        self.tasks_g_id[tasks[cpu]]]
            task = self._scatter(tasks)
            goal = self._scatter(goals)
            self.count += 1
            obs = env.reset_task_goal(goal=goal, task=task, directly=False, eval=self.eval)
        else:
            goals = []
            if self.rank == 0:
                goals = [np.random.uniform(-1, 1, self.dims['g']) for _ in range(self.nb_cpu)]
            obs = env.reset_task_goal(goal=self._scatter(goals))
        self.initial_o[i] = obs['observation']
        self.initial_ag[i] = obs['achieved_goal']
        self.g[i] = obs['desired_goal']
        if self.modular:
            self.task_descr[i] = obs['mask']

    def reset_all_rollouts(self):
        self.goals = [[] for _ in range(self.nb_goals_per_rollout)]
        if self.modular:
            self.tasks = [[] for _ in range(self.nb_goals_per_rollout)]
        for i in range(self.rollout_batch_size):
            self.reset_rollout(i)

    # ------------------------------------------------------------------------------------------------------
    def generate_rollouts(self):
        """rollout.py:178-406.  Returns (episode batch-major, CP, n_episodes)."""
        if self.modular and not self.eval:
            self.exploit = True if np.random.random() < 0.1 else False        # competence is measured without noise
            if self.exploit and self.structure == 'curious':
                self.p = 1 / self.nb_tasks * np.ones([self.nb_tasks])
        elif self.eval:
            self.exploit = True
            if self.modular:
                self.p = 1 / self.nb_tasks * np.ones([self.nb_tasks])
        self.reset_all_rollouts()
        B = self.rollout_batch_size
        o = np.empty((B, self.dims['o']), np.float32)
        ag = np.empty((B, self.dims['ag']), np.float32)
        o[:] = self.initial_o
        ag[:] = self.initial_ag
        obs, achieved_goals, acts, goals, successes = [], [], [], [], []
        info_values = [np.empty((self.T, B, self.dims['info_' + key]), np.float32) for key in self.info_keys]
        Qs, task_descrs, changes = [], [], []
        r_competence = np.zeros(B)
        for t in range(self.T):
            if self.structure == 'task_experts' and self.eval:
                act_output = np.zeros([B, self.dims['u']])
                q_output = np.zeros([B, 1])
                for i in range(B):                    # the expert of the demanded task acts (rollout.py:211-223)
                    tsk = int(np.argmax(self.task_descr[i]))
                    out = self.policy[tsk].get_actions(o[i:i + 1], ag[i:i + 1], self.g[i:i + 1],
                                                       task_descr=self.task_descr[i:i + 1], compute_Q=self.compute_Q,
                                                       noise_eps=0., random_eps=0., use_target_net=self.use_target_net)
                    if self.compute_Q:
                        act_output[i, :], q_output[i, 0] = out[0], np.asarray(out[1]).reshape(-1)[0]
                    else:
                        act_output[i, :] = out
                policy_output = [act_output, q_output] if self.compute_Q else act_output
            else:
                policy_output = self.policy.get_actions(
                    o, ag, self.g, task_descr=self.task_descr if self.modular else None, compute_Q=self.compute_Q,
                    noise_eps=self.noise_eps if not self.exploit else 0.,
                    random_eps=self.random_eps if not self.exploit else 0., use_target_net=self.use_target_net)
            if self.compute_Q:
                u, Q = policy_output
                Qs.append(Q)
            else:
                u = policy_output
            if u.ndim == 1:
                u = u.reshape(1, -1)
            o_new = np.empty((B, self.dims['o']))
            ag_new = np.empty((B, self.dims['ag']))
            success = np.zeros(B)
            for i in range(B):
                curr_o_new, r_competence[i], _, info = self.envs[i].step(u[i])   # reward is recomputed for HER
                if 'is_success' in info:
                    success[i] = info['is_success']
                o_new[i] = curr_o_new['observation']
                ag_new[i] = curr_o_new['achieved_goal']
                self.g[i] = curr_o_new['desired_goal']
                for idx, key in enumerate(self.info_keys):
                    info_values[idx][t, i] = info[key]
            if np.isnan(o_new).any():
                self.reset_all_rollouts()
                return self.generate_rollouts()
            obs.append(o.copy())
            achieved_goals.append(ag.copy())
            successes.append(success.copy())
            acts.append(u.copy())
            goals.append(self.g.copy())
            o[...] = o_new
            ag[...] = ag_new
            if self.modular:
                task_descrs.append(self.task_descr.copy())
                changes.append(np.abs(achieved_goals[0] - ag) > 1e-3)
        obs.append(o.copy())
        achieved_goals.append(ag.copy())
        episode = dict(o=obs, u=acts, g=goals, ag=achieved_goals)
        if self.modular:
            episode['task_descr'] = task_descrs
            episode['change'] = changes
        self.initial_o[:] = o
        for key, value in zip(self.info_keys, info_values):
            episode['info_{}'.format(key)] = value
        successful = np.array(successes)[-1, :]
        assert successful.shape == (B,)
        self.success_history.append(np.mean(successful))
        self.reward_history.append(r_competence.copy())
        if self.compute_Q:
            self.Q_history.append(np.mean(Qs))
        self.n_episodes += B * self.nb_cpu
        if self.modular:
            if self.exploit:
                tasks_c = [int(self.envs[i].unwrapped.task) for i in range(B)]
                succ_c = successful.tolist()
            else:
                tasks_c, succ_c = [], []
            if self.rank == 0:
                self.task_history.extend([t for t in self.tasks if t != []])
                self.goal_history.extend([g for g in self.goals if len(g)])
            self.CP, p = self.tracker.update(tasks_c, succ_c)
            self.C = self.tracker.C
            if not self.eval:
                self.p = np.asarray(p, np.float64).copy()
        return convert_episode_to_batch_major(episode), self.CP, self.n_episodes

    # ------------------------------------------------------------------------------------------------------
    def clear_history(self):
        self.success_history.clear()
        self.reward_history.clear()
        self.Q_history.clear()

    def clear_competence_queue(self):
        self.tracker.clear_competence_queue()

    def current_success_rate(self):
        return np.mean(self.success_history)

    def current_mean_Q(self):
        return np.mean(self.Q_history)

    def save_policy(self, path):
        with open(path, 'wb') as f:
            pickle.dump(self.policy, f)
        try:
            self.policy.save_weights(path)
        except Exception:
            pass

    def logs(self, prefix='worker'):
        logs = [('success_rate', np.mean(self.success_history)), ('avg_reward', np.mean(self.reward_history))]
        if self.compute_Q:
            logs += [('mean_Q', np.mean(self.Q_history))]
        logs += [('episode', self.n_episodes)]
        if prefix != '' and not prefix.endswith('/'):
            return [(prefix + '/' + key, val) for key, val in logs]
        return logs

    def additional_logs(self, prefix='worker'):
        logs = []
        if self.modular:
            Cs = self.get_C()
            for i in range(self.nb_tasks):
                logs += [('C_task' + str(i), '%.3g' % Cs[i])]
                if not self.eval:
                    logs += [('CP_task' + str(i), '%.3g' % self.get_CP()[i])]
                    logs += [('p_task' + str(i), '%.3g' % self.p[i])]
        if prefix != '' and not prefix.endswith('/'):
            return [(prefix + '/' + key, val) for key, val in logs]
        return logs

    def get_CP(self):
        return self.tracker.get_CP()

    def get_C(self):
        return self.tracker.get_C()

    def seed(self, seed):
        for idx, env in enumerate(self.envs):
            env.seed(seed + 1000 * idx)
