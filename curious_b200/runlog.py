"""Run records on disk in the reference's formats, so that its analysis scripts
(baselines/her/analysis/plot.py:29-75, plot_multi.py:49-95, experiment/plot.py:61-75) read runs of this package:

  <dir>/progress.csv   one row per epoch, comma separated, header = keys in the order they first appeared, a key that
                       shows up later adds a column and pads the older rows (baselines/logger.py:101-132)
  <dir>/params.json    the run's parameters as one JSON object (experiment/train.py:255-266)
  <dir>/log.txt        free-text lines (logger.info)
  <dir>/policy_latest.pkl, policy_best.pkl, policy_<epoch>.pkl   pickled policies (train.py:53-55,196-206) - written by
                       RolloutWorker.save_policy, next to DDPG.save_weights' <path>_weights.pkl (ddpg.py:481-497)

Only rank 0 writes these (train.py:51-59,182,264); the other ranks keep their free-text lines in log-rank<NNN>.txt
(logger.py:359-367) and swallow the rest.
"""
import json
import os
import time

import numpy as np


class ProgressCSV(object):
    """progress.csv writer.  Rows are kept in memory (one per epoch) and the file is rewritten when a new key appears,
    which produces the same bytes as the reference's in-place header rewrite + comma padding."""

    def __init__(self, filename, resume=False, keep=None):
        """resume: continue an existing file; keep(row) -> bool drops the rows a resumed run is going to write again."""
        self.filename = filename
        self.columns = []
        self.rows = []
        if resume and os.path.exists(filename):
            lines = open(filename).read().splitlines()
            if lines:
                self.columns = lines[0].split(',')
                for line in lines[1:]:
                    row = {c: (v if v != '' else None) for c, v in zip(self.columns, line.split(','))}
                    if keep is None or keep(row):
                        self.rows.append(row)
        self.file = open(filename, 'w')
        if self.columns:
            self.file.write(','.join(self.columns) + '\n')
            for row in self.rows:
                self.file.write(self._line(self.columns, row))
            self.file.flush()

    @staticmethod
    def _line(columns, row):
        return ','.join('' if row.get(c) is None else str(row[c]) for c in columns) + '\n'

    def writekvs(self, kvs):
        row = dict(kvs)
        fresh = [k for k in row if k not in self.columns]
        self.rows.append(row)
        if fresh:
            self.columns.extend(fresh)
            self.file.seek(0)
            self.file.truncate()
            self.file.write(','.join(self.columns) + '\n')
            for old in self.rows[:-1]:
                self.file.write(self._line(self.columns, old))
        self.file.write(self._line(self.columns, row))
        self.file.flush()

    def close(self):
        self.file.close()


class RunLog(object):
    """The slice of baselines.logger the HER driver uses: get_dir / record_tabular / dump_tabular / info
    (logger.py:192-245), plus params.json."""

    def __init__(self, directory, rank=0, echo=False, resume_after_epoch=None):
        """resume_after_epoch: continue the files of an earlier run, keeping its rows up to that epoch."""
        self.dir = directory
        self.active = rank == 0 and directory is not None
        self.echo = echo
        self.row = {}
        self.csv = self.txt = None
        if self.active:
            os.makedirs(directory, exist_ok=True)
            resume = resume_after_epoch is not None
            self.csv = ProgressCSV(os.path.join(directory, 'progress.csv'), resume=resume,
                                   keep=(lambda row: int(float(row.get('epoch') or 0)) <= resume_after_epoch) if resume else None)
            self.txt = open(os.path.join(directory, 'log.txt'), 'a' if resume else 'w')
        elif directory is not None:
            # the other ranks keep their free-text lines in log-rank<NNN>.txt (logger.py:359-367)
            os.makedirs(directory, exist_ok=True)
            self.txt = open(os.path.join(directory, 'log-rank%03i.txt' % rank),
                            'a' if resume_after_epoch is not None else 'w')
        self.t0 = time.time()

    def get_dir(self):
        return self.dir

    def record_tabular(self, key, val):
        self.row[key] = val

    def dump_tabular(self):
        if self.active:
            self.csv.writekvs(self.row)
            if self.echo:
                width = max(len(k) for k in self.row) if self.row else 0
                print('\n'.join('| %-*s | %s' % (width, k, v) for k, v in self.row.items()), flush=True)
        self.row = {}

    def info(self, *args):
        if self.txt is not None:
            line = ' '.join(str(a) for a in args)
            self.txt.write(line + '\n')
            self.txt.flush()
            if self.echo and self.active:
                print(line, flush=True)

    def write_params(self, params):
        """params.json; values that JSON cannot hold (callables, arrays) are written as their repr / lists."""
        if not self.active:
            return

        def plain(v):
            if isinstance(v, (np.integer,)):
                return int(v)
            if isinstance(v, (np.floating,)):
                return float(v)
            if isinstance(v, np.ndarray):
                return v.tolist()
            return repr(v)
        with open(os.path.join(self.dir, 'params.json'), 'w') as f:
            json.dump(params, f, default=plain)

    def close(self):
        if self.csv is not None:
            self.csv.close()
        if self.txt is not None:
            self.txt.close()


def mpi_average(value, comm=None):
    """Mean over ranks of the mean of `value` (her/util.py:141-146 -> mpi_moments): a scalar or a list per rank."""
    from .parallel import world
    if isinstance(value, (list, tuple, np.ndarray)):
        v = np.asarray(value, np.float64).reshape(-1)
        if v.size == 0:
            v = np.zeros(1)
    else:
        v = np.array([float(value)])
    group, n = world(comm)
    if n <= 1:
        return float(v.mean())
    import torch
    import torch.distributed as dist
    # pooled mean (mpi_moments sums values and counts over ranks, mpi_moments.py:6-17)
    acc = torch.tensor([v.sum(), float(v.size)], dtype=torch.float64)
    if dist.get_backend(group) == 'nccl':
        acc = acc.cuda()
    dist.all_reduce(acc, group=group)
    return float(acc[0].item() / acc[1].item())
