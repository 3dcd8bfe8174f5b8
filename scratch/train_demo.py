import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
from curious_b200.train import make_experiment, train
np.random.seed(0)
structure = os.environ.get('STRUCT', 'curious')
replay = {'curious': 'replay_task_cp_buffer', 'task_experts': 'replay_current_task_buffer', 'flat': ''}[structure]
exp = make_experiment(nb_tasks=int(os.environ.get('N', 4)), n_controllable=int(os.environ.get('NC', 3)), structure=structure,
                      task_replay=replay, buffer_size=100000, n_cycles=int(os.environ.get('CYC', 10)), n_batches=40, n_test_rollouts=10)
t0 = time.time()
def log(r):
    print('epoch %d  %.1fs  test success %.2f  C %s  CP %s  p %s' % (
        r['epoch'], time.time() - t0, r['test_success_rate'], np.round(r.get('C', []), 2), np.round(r.get('CP', []), 3), np.round(r.get('p', []), 2)), flush=True)
train(n_epochs=int(os.environ.get('EPOCHS', 12)), log=log, **exp)
