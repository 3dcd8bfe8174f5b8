"""Generate tests/golden/progress_rows.json + progress_golden.csv from the UNMODIFIED reference baselines/logger.py
(CSVOutputFormat, logger.py:101-132) - test infrastructure; build container only (needs /root/reference).

The reference orders the keys of one row by iterating a set difference, i.e. by string hash: with more than one NEW key
in a row the column order depends on PYTHONHASHSEED.  The rows below therefore introduce at most one new key at a time,
which makes the reference's bytes deterministic and comparable (readers go by column name, so the order of several
simultaneous new columns is not part of the format).
"""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'

ROWS = [
    {'epoch': 0},
    {'epoch': 1, 'test/success_rate': '0.25'},
    {'epoch': 2, 'test/success_rate': '0.5', 'train/episode': 100},
    {'epoch': 3, 'train/episode': 150},                                   # a known key missing: empty cell
    {'epoch': 4, 'test/success_rate': '1', 'train/episode': 200, 'train/CP_task0': '0.0125'},
    {'epoch': 5, 'test/success_rate': '1', 'train/episode': 250, 'train/CP_task0': '0', 'Time': 12.5},
]


def main(outdir=None):
    outdir = outdir or os.path.join(ROOT, 'tests', 'golden')
    spec = importlib.util.spec_from_file_location('ref_logger', os.path.join(REF, 'baselines/logger.py'))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = os.path.join(outdir, 'progress_golden.csv')
    w = ref.CSVOutputFormat(out)
    for row in ROWS:
        w.writekvs(dict(row))
    w.close()
    with open(os.path.join(outdir, 'progress_rows.json'), 'w') as f:
        json.dump(ROWS, f)
    print(open(out).read())


if __name__ == '__main__':
    main()
