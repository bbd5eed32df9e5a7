"""HEADS registry: the drop-in boundary of this package.

If mmdet is importable, CPRHead/P2PHead register into mmdet's own `HEADS` (= MODELS) registry with `force=True`, so a
reference config (`bbox_head=dict(type='CPRHead', ...)`) builds this implementation unchanged.  Otherwise a minimal
registry with the same `register_module` / `build` surface is used (mmcv/mmdet cannot be installed offline here).
"""


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self.module_dict[key] = cls
            return cls
        return _reg(module) if module is not None else _reg

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        args = dict(cfg)
        for k, v in (default_args or {}).items():
            args.setdefault(k, v)
        t = args.pop('type')
        cls = self.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f'{t} is not in the {self.name} registry')
        return cls(**args)


try:   # pragma: no cover - mmdet is not installable in the offline build container
    from mmdet.models.builder import HEADS as _MMDET_HEADS
    HEADS = _MMDET_HEADS
    USING_MMDET = True
except Exception:
    HEADS = Registry('head')
    USING_MMDET = False


def register_head(cls):
    HEADS.register_module(name=cls.__name__, force=True, module=cls)
    return cls


def build_head(cfg, default_args=None):
    return HEADS.build(cfg, default_args)


class CfgNode(dict):
    """attribute-style dict (stand-in for mmcv.Config nodes in train_cfg / test_cfg)."""
    __getattr__ = dict.get
