"""Attribute-style config dict + a small YAML loader with ``${a.b}`` interpolation and ``key=value``
overrides — the subset of omegaconf/hydra the reference's trainers use."""
from __future__ import annotations

import re
from typing import Any, Iterable

import yaml


class DictConfig(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = value


def to_config(obj: Any) -> Any:
    if isinstance(obj, dict):
        return DictConfig({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_config(v) for v in obj]
    return obj


_REF = re.compile(r"\$\{([^}]+)\}")


def _lookup(root: dict, path: str):
    cur = root
    for part in path.split("."):
        cur = cur[part]
    return cur


def _resolve(node, root, depth=0):
    if depth > 32:
        raise ValueError("config interpolation does not terminate")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _REF.fullmatch(node)
        if m:  # whole-value reference keeps the referenced type
            return _resolve(_lookup(root, m.group(1)), root, depth + 1)
        if _REF.search(node):
            return _REF.sub(lambda mm: str(_resolve(_lookup(root, mm.group(1)), root, depth + 1)), node)
    return node


def load_config(path: str, overrides: Iterable[str] = ()) -> DictConfig:
    with open(path, "r", encoding="utf-8") as fh:
        raw = yaml.safe_load(fh)
    for ov in overrides:
        key, _, val = ov.partition("=")
        cur = raw
        parts = key.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = yaml.safe_load(val)
    return to_config(_resolve(raw, raw))
