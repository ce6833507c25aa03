"""Config / CLI surface of the reference (utils/VidToMe/config_utils.py:6-73) without omegaconf:
same YAML keys (configs/tclight_default.yaml), recursive ``base_config`` chain with deep merge,
``${a.b}`` interpolation, and the quick-use CLI flags ``--config --base_config -i -p -n --multi_axis``.
"""
from __future__ import annotations

import argparse
import copy
import os
import re
from datetime import datetime
from typing import Any

import yaml


class Config(dict):
    """Nested dict with attribute access (the subset of DictConfig the pipeline uses)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __delattr__(self, k):
        del self[k]

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v: Any) -> Any:
    if isinstance(v, dict) and not isinstance(v, Config):
        return Config({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    return v


def merge(base: dict, over: dict) -> Config:
    """OmegaConf.merge(base, over): deep, `over` wins."""
    out = Config({k: _wrap(copy.deepcopy(v)) for k, v in base.items()})
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = merge(out[k], v)
        else:
            out[k] = _wrap(copy.deepcopy(v))
    return out


_INTERP = re.compile(r"\$\{([^}]+)\}")


def resolve(cfg: Config) -> Config:
    """OmegaConf.resolve: substitute ``${dotted.path}`` references (strings only)."""

    def lookup(path: str):
        cur: Any = cfg
        for p in path.split("."):
            cur = cur[p]
        return cur

    def rec(node):
        if isinstance(node, dict):
            for k in list(node.keys()):
                node[k] = rec(node[k])
            return node
        if isinstance(node, list):
            return [rec(x) for x in node]
        if isinstance(node, str):
            for _ in range(8):
                m = _INTERP.fullmatch(node)
                if m:
                    node = rec(lookup(m.group(1)))
                    if not isinstance(node, str):
                        return node
                    continue
                if not _INTERP.search(node):
                    break
                node = _INTERP.sub(lambda mm: str(rec(lookup(mm.group(1)))), node)
            return node
        return node

    return rec(cfg)


def load_yaml(path: str) -> Config:
    with open(path, "r") as f:
        return _wrap(yaml.safe_load(f) or {})


def load_config_file(config_path: str, base_config: str = None) -> Config:
    """The base_config merge loop of config_utils.py:26-37."""
    config = load_yaml(config_path)
    cur_path, cur = config_path, config
    if base_config is not None:
        cur["base_config"] = base_config
    while "base_config" in cur and cur["base_config"] != cur_path:
        base = load_yaml(cur["base_config"])
        config = merge(base, config)
        cur_path, cur = cur["base_config"], base
    return config


def load_config(print_config: bool = True, argv=None) -> Config:
    """reference config_utils.py:6-65."""
    p = argparse.ArgumentParser()
    p.add_argument("--config", type=str, default="configs/tclight_default.yaml")
    p.add_argument("--base_config", type=str, default=None)
    p.add_argument("--input_path", "-i", type=str, default=None)
    p.add_argument("--prompt", "-p", type=str, default=None)
    p.add_argument("--negative_prompt", "-n", type=str, default=None)
    p.add_argument("--multi_axis", action="store_true")
    args = p.parse_args(argv)
    config = load_config_file(args.config, args.base_config)
    if args.input_path is not None and str(config.data.scene_type).lower() == "video":
        config.data.rgb_path = args.input_path
    if args.multi_axis:
        config.generation.alpha_t = 0.01
    if args.negative_prompt is not None:
        config.generation.negative_prompt = args.negative_prompt
    if args.prompt is not None or isinstance(config.generation.prompt, str):
        args.prompt = config.generation.prompt if args.prompt is None else args.prompt
        date_time = datetime.now().strftime("%m-%d-%Y")
        video_name = os.path.splitext(os.path.basename(config.data.rgb_path))[0]
        config.work_dir = os.path.join(config.work_dir, date_time, video_name)
        os.makedirs(config.work_dir, exist_ok=True)
        existing = os.listdir(config.work_dir)
        save_idx = max([int(x[-5:]) for x in existing]) + 1 if existing else 0
        config.generation.prompt = Config({f"{args.prompt}-{str(save_idx).zfill(5)}": args.prompt})
    prompt = config.generation.prompt
    if isinstance(prompt, str):
        prompt = Config({"edit": prompt})
    config.generation.prompt = prompt
    config = resolve(config)
    if print_config:
        print("[INFO] loaded config:")
        print(yaml.safe_dump(to_plain(config), sort_keys=False))
    return config


def to_plain(node):
    if isinstance(node, dict):
        return {k: to_plain(v) for k, v in node.items()}
    if isinstance(node, list):
        return [to_plain(v) for v in node]
    return node


def save_config(config: Config, path: str, gene: bool = False, inv: bool = False):
    """reference config_utils.py:67-73."""
    os.makedirs(path, exist_ok=True)
    cfg = to_plain(config)
    if gene:
        cfg.pop("inversion", None)
    if inv:
        cfg.pop("generation", None)
    with open(os.path.join(path, "config.yaml"), "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)


DEFAULT_YAML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "tclight_default.yaml")


def default_config(**generation_overrides) -> Config:
    """The shipped defaults (same keys/values as the reference's configs/tclight_default.yaml)."""
    cfg = resolve(load_yaml(DEFAULT_YAML))
    for k, v in generation_overrides.items():
        cfg.generation[k] = v
    if cfg.model_key is None:
        cfg.model_key = "iclight"
    return cfg
