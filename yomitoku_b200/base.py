"""BaseModule / BaseModelCatalog / observer: the reference's module protocol for the hot path.

Mirrors reference src/yomitoku/base.py:36-142 (same names, argument meaning and error behaviour):
catalog lookup is case-insensitive and raises ValueError on unknown / duplicate names; `observer` wraps `__call__`
with wall-clock logging and re-raises; `device` is a torch.device with the reference's CUDA->CPU warning fallback for
the *plumbing* only - the CUDA models themselves refuse to run without a GPU (no CPU fallback on the hot path).
"""
import logging
import time

import torch
import yaml

from .config import load_config

_loggers = {}


def set_logger(name, level="INFO"):
    if name in _loggers:
        return _loggers[name]
    logger = logging.getLogger(name)
    logger.setLevel(level)
    handler = logging.StreamHandler()
    handler.setLevel(level)
    handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
    logger.addHandler(handler)
    _loggers[name] = logger
    return logger


logger = set_logger(__name__, "WARNING")


def observer(cls, func):
    def wrapper(*args, **kwargs):
        try:
            start = time.time()
            result = func(*args, **kwargs)
            elapsed = time.time() - start
            logger.info(f"{cls.__name__} {func.__name__} elapsed_time: {elapsed}")
        except Exception as e:
            logger.error(f"Error occurred in {cls.__name__} {func.__name__}: {e}")
            raise e
        return result

    wrapper.__wrapped__ = getattr(func, "__wrapped__", func)
    return wrapper


class BaseModelCatalog:
    def __init__(self):
        self.catalog = {}

    def get(self, model_name):
        model_name = model_name.lower()
        if model_name in self.catalog:
            return self.catalog[model_name]
        raise ValueError(f"Unknown model: {model_name}")

    def register(self, model_name, config, model):
        if model_name in self.catalog:
            raise ValueError(f"{model_name} is already registered.")
        self.catalog[model_name] = (config, model)

    def list_model(self):
        return list(self.catalog.keys())


class BaseModule:
    model_catalog = None

    def __init__(self):
        if self.model_catalog is None:
            raise NotImplementedError
        if not issubclass(self.model_catalog.__class__, BaseModelCatalog):
            raise ValueError(f"{self.model_catalog.__class__} is not SubClass BaseModelCatalog.")
        if len(self.model_catalog.list_model()) == 0:
            raise ValueError("No model is registered.")

    def __new__(cls, *args, **kwds):
        logger.info(f"Initialize {cls.__name__}")
        # the reference re-wraps on every instantiation; wrap once per class instead
        call = cls.__dict__.get("__call__") or cls.__call__
        if not getattr(call, "_ytk_observed", False):
            wrapped = observer(cls, cls.__call__)
            wrapped._ytk_observed = True
            cls.__call__ = wrapped
        return super().__new__(cls)

    def load_model(self, name, path_cfg, from_pretrained=True):
        default_cfg, Net = self.model_catalog.get(name)
        self._cfg = load_config(default_cfg, path_cfg)
        if from_pretrained:
            self.model = Net.from_pretrained(self._cfg.hf_hub_repo, cfg=self._cfg)
        else:
            self.model = Net(cfg=self._cfg)

    def save_config(self, path_cfg):
        with open(path_cfg, "w", encoding="utf-8") as f:
            yaml.safe_dump(_plain(self._cfg), f, allow_unicode=True)

    def log_config(self):
        logger.info(yaml.safe_dump(_plain(self._cfg), allow_unicode=True))

    @classmethod
    def catalog(cls):
        display = ""
        for model in cls.model_catalog.list_model():
            display += f"{model} "
        logger.info(f"{cls.__name__} Implemented Models")
        logger.info(display)

    @property
    def device(self):
        return self._device

    @device.setter
    def device(self, device):
        device = str(device)
        if "cuda" in device:
            if torch.cuda.is_available():
                self._device = torch.device(device)
            else:
                self._device = torch.device("cpu")
                logger.warning("CUDA is not available. Use CPU instead.")
        elif "mps" in device:
            self._device = torch.device("cpu")
            logger.warning("MPS is not available. Use CPU instead.")
        else:
            self._device = torch.device("cpu")


def _plain(cfg):
    if isinstance(cfg, dict):
        return {k: _plain(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [_plain(v) for v in cfg]
    return cfg
