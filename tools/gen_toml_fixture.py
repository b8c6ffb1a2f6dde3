#!/usr/bin/env python
"""tests/golden/flylight_default_kwargs.json = the keyword arguments run_ppp.py hands
to the assembly stage for the shipped flylight setup, taken VERBATIM from
/root/reference/experiments/flylight/setups/setup01/default.toml:
[vote_instances] + [model] (+ [visualize] and the [prediction] keys for the blockwise
call, run_ppp.py:1163-1190).  Runs only where the reference is present."""
import json
import os
import sys
import tomllib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOML = '/root/reference/experiments/flylight/setups/setup01/default.toml'


def load():
    with open(TOML, 'rb') as f:
        cfg = tomllib.load(f)
    pred = cfg.get('prediction', {})
    return dict(vote_instances=cfg['vote_instances'], model=cfg['model'],
                visualize=cfg.get('visualize', {}),
                prediction={k: pred.get(k) for k in ('aff_key', 'numinst_key', 'fg_key',
                                                     'fg_folder', 'fg_thresh', 'output_format')})


if __name__ == '__main__':
    out = os.path.join(ROOT, 'tests', 'golden', 'flylight_default_kwargs.json')
    with open(out, 'w') as f:
        json.dump(load(), f, indent=1, sort_keys=True)
    print(out)
    sys.exit(0)
