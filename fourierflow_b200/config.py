"""Config loader for the reference's experiment YAML files, without hydra / omegaconf / pytorch_lightning.

The reference builds everything from one YAML per experiment (`experiments/**/config.yaml`) through
``hydra.compose`` + ``hydra.utils.instantiate`` (fourierflow/commands/train.py:38-40,66-67) with three custom
OmegaConf resolvers registered in fourierflow/__init__.py:20-24 (``get_method``, ``import``, ``eval``) plus the
built-in ``oc.env``.  None of those packages exist on the GPU box, so this module restates the small subset the
experiment files use:

* ``${oc.env:VAR}`` / ``${oc.env:VAR,default}``, ``${get_method: pkg.mod.fn}``, ``${import: pkg.mod.NAME}``,
  ``${eval: expr}`` and plain node references ``${a.b.c}``;
* recursive ``_target_`` instantiation with ``_args_`` / ``_partial_`` and keyword overrides.

Targets under ``fourierflow.modules`` / ``fourierflow.routines`` are mapped onto this package's CUDA-backed mirrors,
so the ``routine:`` block of e.g. ``experiments/torus_li/markov/24_layers/config.yaml`` instantiates unchanged.
Anything else from the reference tree (builders, callbacks, schedulers: training harness) is
outside the hot path: such a target resolves to a :class:`MissingTarget` placeholder that raises when *used*, so
that a config still loads and its in-scope parts still run.
"""
from __future__ import annotations

import importlib
import os
import re
from typing import Any, Callable, Dict, Mapping, Optional, Sequence

import yaml

#: reference package prefix -> package that provides the same names here
TARGET_MAP = {
    "fourierflow.modules": "fourierflow_b200.modules",
    "fourierflow.routines": "fourierflow_b200.routines",
}


class MissingTarget:
    """Stands in for a reference symbol that is not part of the B200 hot path (or is not installed here)."""

    def __init__(self, path: str, why: str):
        self.path, self.why = path, why

    def __call__(self, *a, **k):
        raise RuntimeError(f"config target '{self.path}' is not available in fourierflow_b200: {self.why}")

    def __repr__(self):
        return f"MissingTarget({self.path!r})"


def _map_path(path: str) -> str:
    for src, dst in TARGET_MAP.items():
        if path == src or path.startswith(src + "."):
            return dst + path[len(src):]
    return path


def locate(path: str) -> Any:
    """Import ``pkg.mod.attr`` (the job of hydra.utils.get_method / get_class and fourierflow.utils.import_string)."""
    path = path.strip()
    mapped = _map_path(path)
    parts = mapped.split(".")
    err: Optional[Exception] = None
    for i in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module(".".join(parts[:i]))
        except ImportError as e:          # keep looking for a shorter module prefix
            err = err or e
            continue
        try:
            for name in parts[i:]:
                obj = getattr(obj, name)
            return obj
        except AttributeError as e:
            err = e
            break
    if path.startswith("fourierflow.") or path.startswith("pytorch_lightning.") or path.startswith("wandb"):
        return MissingTarget(path, f"outside the F-FNO forward hot path ({err})")
    raise ImportError(f"cannot locate '{path}': {err}")


# ---------------------------------------------------------------------------------------------------------
# interpolation
# ---------------------------------------------------------------------------------------------------------
_INTERP = re.compile(r"\$\{([^${}]*)\}")


def _resolver_env(arg: str) -> str:
    name, _, default = arg.partition(",")
    name = name.strip()
    if name in os.environ:
        return os.environ[name]
    if default != "" or "," in arg:
        return default.strip()
    raise KeyError(f"environment variable '{name}' is not set (config uses ${{oc.env:{name}}})")


RESOLVERS: Dict[str, Callable[[str], Any]] = {
    "oc.env": _resolver_env,
    "get_method": locate,
    "import": locate,
    "eval": lambda expr: eval(expr, {"__builtins__": {}}, {}),   # the reference registers plain eval (:24)
}


def _select(root: Mapping, dotted: str) -> Any:
    node: Any = root
    for key in dotted.split("."):
        node = node[int(key)] if isinstance(node, list) else node[key]
    return node


def _resolve_str(text: str, root: Mapping, depth: int = 0) -> Any:
    if depth > 16:
        raise ValueError(f"interpolation cycle in '{text}'")

    def one(expr: str) -> Any:
        name, sep, arg = expr.partition(":")
        if sep and name.strip() in RESOLVERS:
            return RESOLVERS[name.strip()](arg.strip())
        return _resolve(_select(root, expr.strip()), root, depth + 1)

    # innermost interpolations first (``${eval:2 * ${import:numpy.pi}}``); when what is left is exactly one
    # interpolation its resolved value keeps its type (float, callable, ...), otherwise values are spliced as text
    out = text
    while True:
        m = _INTERP.fullmatch(out.strip())
        if m:
            return one(m.group(1))
        m = _INTERP.search(out)
        if not m:
            return out
        val = one(m.group(1))
        out = out[:m.start()] + (repr(val) if isinstance(val, float) else str(val)) + out[m.end():]


def _resolve(node: Any, root: Mapping, depth: int = 0, strict: bool = True) -> Any:
    if isinstance(node, str) and "${" in node:
        try:
            return _resolve_str(node, root, depth)
        except KeyError:
            if strict:
                raise
            return node                      # e.g. ${oc.env:DATA_ROOT} of the data builder on a box without data
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth, strict) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth, strict) for v in node]
    return node


def _apply_overrides(cfg: dict, overrides: Sequence[str]) -> None:
    """``a.b.c=value`` assignments, the form hydra's CLI overrides take (commands/train.py:38-40)."""
    for ov in overrides:
        key, sep, val = ov.partition("=")
        if not sep:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        node = cfg
        parts = key.lstrip("+").split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {}) if isinstance(node, dict) else node[int(p)]
        node[parts[-1]] = yaml.safe_load(val)


def load_config(source: str, overrides: Sequence[str] = (), resolve: bool = True, strict: bool = False) -> dict:
    """Parse an experiment YAML (a path or the YAML text itself), apply overrides, resolve interpolations.
    OmegaConf resolves lazily, on access; here everything is resolved up front, so by default (``strict=False``) a
    value whose environment variable / reference is missing keeps its ``${...}`` text instead of failing the load."""
    text = open(source).read() if "\n" not in source and os.path.exists(source) else source
    cfg = yaml.safe_load(text) or {}
    _apply_overrides(cfg, overrides)
    return _resolve(cfg, cfg, strict=strict) if resolve else cfg


# ---------------------------------------------------------------------------------------------------------
# instantiation
# ---------------------------------------------------------------------------------------------------------
def instantiate(node: Any, *args, **overrides) -> Any:
    """Recursive ``_target_`` instantiation (hydra.utils.instantiate semantics for the keys the configs use:
    ``_target_``, ``_args_``, ``_partial_``; nested nodes are instantiated first; ``overrides`` replace keywords)."""
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "_target_" not in node:
        return {k: instantiate(v) for k, v in node.items()}
    target = node["_target_"]
    fn = locate(target) if isinstance(target, str) else target
    pos = [instantiate(a) for a in node.get("_args_", [])] + list(args)
    kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_args_", "_partial_")}
    kwargs.update(overrides)
    if node.get("_partial_", False):
        import functools
        return functools.partial(fn, *pos, **kwargs)
    return fn(*pos, **kwargs)


def load_routine(source: str, overrides: Sequence[str] = ()):
    """The ``routine:`` block of an experiment config as a ready module tree (what commands/train.py:67 builds),
    backed by libffno_b200.  The builder / trainer / callbacks blocks are left untouched in the returned config."""
    raw = load_config(source, overrides, resolve=False)
    if "routine" not in raw:
        raise KeyError("config has no 'routine' block")
    routine = instantiate(_resolve(raw["routine"], raw, strict=True))
    return routine, _resolve(raw, raw, strict=False)
