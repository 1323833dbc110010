"""Module protocol of the hot path: model catalogs, the per-module base class and its call instrumentation.

Behavioural contract taken from reference src/yomitoku/base.py:36-142: `BaseModelCatalog`
(`register` / `get` / `list_model`; lookup is case-insensitive, unknown and duplicate names raise ValueError),
`BaseModule` (class attribute `model_catalog`, `load_model(name, path_cfg, from_pretrained)`, `save_config`,
`log_config`, `catalog()`, the `device` property) and `observer` (every module's `__call__` is timed and logged,
exceptions are logged and re-raised).  `device` accepts "cuda..." and falls back to CPU with a warning when no GPU is
visible - for the *plumbing* only: the CUDA models themselves refuse to run without a GPU (no CPU fallback on the hot
path).
"""
import functools
import logging
import time

import torch
import yaml

from .config import load_config


def set_logger(name, level="INFO"):
    """One stream handler per logger name, created on first use."""
    log = logging.getLogger(name)
    if not getattr(log, "_ytk_configured", False):
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        handler.setLevel(level)
        log.addHandler(handler)
        log.setLevel(level)
        log._ytk_configured = True
    return log


logger = set_logger(__name__, "WARNING")


def observer(cls, func):
    """Wraps a module's `__call__`: wall-clock time at INFO, failures at ERROR (and re-raised)."""
    label = "%s %s" % (cls.__name__, getattr(func, "__name__", "call"))

    @functools.wraps(func)
    def timed(*args, **kwargs):
        t0 = time.perf_counter()
        try:
            out = func(*args, **kwargs)
        except Exception as exc:
            logger.error("Error occurred in %s: %s", label, exc)
            raise
        logger.info("%s elapsed_time: %s", label, time.perf_counter() - t0)
        return out

    timed._ytk_observed = True
    return timed


class BaseModelCatalog:
    """name -> (config dataclass, network class); names are matched case-insensitively on lookup."""

    def __init__(self):
        self.catalog = {}

    def register(self, model_name, config, model):
        if model_name in self.catalog:
            raise ValueError(f"{model_name} is already registered.")
        self.catalog[model_name] = (config, model)

    def get(self, model_name):
        try:
            return self.catalog[model_name.lower()]
        except KeyError:
            raise ValueError(f"Unknown model: {model_name.lower()}") from None

    def list_model(self):
        return [*self.catalog]


def _to_builtin(node):
    """Config tree -> plain dicts / lists for YAML output."""
    if isinstance(node, dict):
        return {key: _to_builtin(val) for key, val in node.items()}
    if isinstance(node, (list, tuple)):
        return [_to_builtin(val) for val in node]
    return node


class BaseModule:
    model_catalog = None

    def __new__(cls, *args, **kwargs):
        logger.info("Initialize %s", cls.__name__)
        # instrument __call__ once per class (the reference re-wraps it on every instantiation)
        if not getattr(cls.__call__, "_ytk_observed", False):
            cls.__call__ = observer(cls, cls.__call__)
        return super().__new__(cls)

    def __init__(self):
        catalog = self.model_catalog
        if catalog is None:
            raise NotImplementedError
        if not isinstance(catalog, BaseModelCatalog):
            raise ValueError(f"{catalog.__class__} is not SubClass BaseModelCatalog.")
        if not catalog.list_model():
            raise ValueError("No model is registered.")

    # ------------------------------------------------------------------------------------------ model / config
    def load_model(self, name, path_cfg, from_pretrained=True):
        cfg_class, net_class = self.model_catalog.get(name)
        self._cfg = load_config(cfg_class, path_cfg)
        self.model = (net_class.from_pretrained(self._cfg.hf_hub_repo, cfg=self._cfg) if from_pretrained
                      else net_class(cfg=self._cfg))

    def _cfg_yaml(self):
        return yaml.safe_dump(_to_builtin(self._cfg), allow_unicode=True)

    def save_config(self, path_cfg):
        with open(path_cfg, "w", encoding="utf-8") as fh:
            fh.write(self._cfg_yaml())

    def log_config(self):
        logger.info(self._cfg_yaml())

    @classmethod
    def catalog(cls):
        logger.info("%s Implemented Models", cls.__name__)
        logger.info("".join(name + " " for name in cls.model_catalog.list_model()))

    # ------------------------------------------------------------------------------------------ device
    @property
    def device(self):
        return self._device

    @device.setter
    def device(self, device):
        wanted = str(device)
        self._device = torch.device("cpu")
        if "cuda" in wanted:
            if torch.cuda.is_available():
                self._device = torch.device(wanted)
            else:
                logger.warning("CUDA is not available. Use CPU instead.")
        elif "mps" in wanted:
            logger.warning("MPS is not available. Use CPU instead.")
