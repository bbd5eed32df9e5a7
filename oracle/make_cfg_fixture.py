"""Extracts the head configs of every shipped CPR / P2P config of the reference into tests/golden/reference_head_cfgs.json.
(test infrastructure; run in the build container: python -m oracle.make_cfg_fixture)

A minimal re-implementation of mmcv.Config's `_base_` inheritance (dict merge, `_delete_` support) is enough for these files.
"""
import glob
import json
import os
import runpy

REF = '/root/reference/TOV_mmdetection'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_head_cfgs.json')


def merge(a, b):
    out = dict(a)
    for k, v in b.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get('_delete_', False):
            out[k] = merge(out[k], v)
        else:
            out[k] = {kk: vv for kk, vv in v.items() if kk != '_delete_'} if isinstance(v, dict) else v
    return out


def load(path):
    ns = runpy.run_path(path)
    cfg = {k: v for k, v in ns.items() if not k.startswith('__') and not callable(v) and not isinstance(v, type(os))}
    base = cfg.pop('_base_', [])
    base = [base] if isinstance(base, str) else base
    merged = {}
    for b in base:
        merged = merge(merged, load(os.path.normpath(os.path.join(os.path.dirname(path), b))))
    return merge(merged, cfg)


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (int, float, str, bool)) or o is None:
        return o
    return repr(o)


def main():
    out = {}
    files = sorted(glob.glob(REF + '/configs2/*/coarsepointv2/*.py') + glob.glob(REF + '/configs2/*/p2p/*.py') +
                   glob.glob(REF + '/configs2/*/*/p2p/*.py'))
    for f in files:
        try:
            cfg = load(f)
        except Exception as e:
            out[os.path.relpath(f, REF)] = dict(error=repr(e)[:200])
            continue
        m = cfg.get('model', {})
        head = m.get('bbox_head')
        if not isinstance(head, dict):
            continue
        out[os.path.relpath(f, REF)] = jsonable(dict(bbox_head=head, train_cfg=m.get('train_cfg'), test_cfg=m.get('test_cfg')))
    json.dump(out, open(OUT, 'w'), indent=1, sort_keys=True)
    print(len(out), 'configs ->', OUT)
    for k, v in out.items():
        print(' ', k, v.get('error') or v['bbox_head'].get('type'))


if __name__ == '__main__':
    main()
