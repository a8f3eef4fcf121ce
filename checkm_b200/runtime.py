"""Process-wide engine and model-database cache.

One engine per process per GPU (include/ckm.h); model databases are parsed, configured and uploaded once and kept
resident, keyed by path + mtime.  The device index comes from CKM_DEVICE, else LOCAL_RANK (torchrun), else 0."""
import os

from .engine import Engine

_engine = None
_extra = []
_models = {}


def device_index():
    for var in ('CKM_DEVICE', 'LOCAL_RANK'):
        v = os.environ.get(var)
        if v is not None and v != '':
            return int(v)
    return 0


def engine():
    global _engine
    if _engine is None:
        _engine = Engine(device_index())
    return _engine


def engines(n):
    """`n` engines on this process's GPU for pipelined batches: engine 0 is the process-wide one, the others are extra
    streams + workspaces on the same device.  One host thread drives one engine at a time (include/ckm.h); model
    databases are device-resident and shared by all of them."""
    global _extra
    first = engine()
    while len(_extra) < n - 1:
        _extra.append(Engine(first.device))
    return [first] + _extra[:n - 1]


def models_for(path):
    key = os.path.abspath(path)
    stamp = os.path.getmtime(path)
    hit = _models.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    if hit is not None:
        hit[1].close()
    m = engine().load_models(path)
    _models[key] = (stamp, m)
    return m


def shutdown():
    global _engine
    for _, m in _models.values():
        m.close()
    _models.clear()
    while _extra:
        _extra.pop().close()
    if _engine is not None:
        _engine.close()
        _engine = None
