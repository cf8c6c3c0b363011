"""`Model` base: a plain `torch.nn.Module` that records its constructor arguments.

The reference's `Model` (src/graphnet/models/model.py:21-108) is
Logger + Configurable + LightningModule with a metaclass that stores the
constructor arguments of every instance so it can be rebuilt from a config
(src/graphnet/utilities/config/model_config.py:317-352). The hot path needs
only that contract (keyword-reconstructible, `state_dict` key names, the
`_gnn` -> `backbone` rename on load), not Lightning or pydantic.
"""

from __future__ import annotations

import inspect
from abc import ABCMeta
from typing import Any, Dict, Union

import torch


class _ConfigSaver(ABCMeta):
    """Record `cls(*args, **kwargs)` as `{class_name, arguments}` on the instance."""

    def __call__(cls, *args: Any, **kwargs: Any) -> Any:
        obj = super().__call__(*args, **kwargs)
        try:
            bound = inspect.signature(cls.__init__).bind(obj, *args, **kwargs)
            bound.apply_defaults()
            arguments = {k: v for k, v in bound.arguments.items() if k != "self"}
            arguments.update(arguments.pop("kwargs", {}) or {})
        except TypeError:  # pragma: no cover
            arguments = dict(kwargs)
        object.__setattr__(obj, "_config", {"class_name": cls.__name__, "arguments": arguments})
        return obj


class Model(torch.nn.Module, metaclass=_ConfigSaver):
    """Base class of the re-implemented `graphnet.models` components."""

    def __init__(self, name: Any = None, class_name: Any = None, **_: Any) -> None:
        super().__init__()

    @property
    def config(self) -> Dict[str, Any]:
        return self._config

    @classmethod
    def from_config(cls, source: Dict[str, Any], **_: Any) -> "Model":
        """Rebuild from `{class_name, arguments}` (cf. model.py:81-108)."""
        # the reference's YAML files nest every model as {"ModelConfig": {class_name, arguments}} (configs/models/*.yml,
        # utilities/config/model_config.py:317-346); both forms are accepted
        if isinstance(source, dict) and set(source) == {"ModelConfig"}:
            source = source["ModelConfig"]
        registry = {c.__name__: c for c in _all_subclasses(Model)}
        if source["class_name"] not in registry:
            raise KeyError(f"Model.from_config: class {source['class_name']!r} is not part of the DynEdge hot path "
                           f"(known: {sorted(registry)})")
        klass = registry[source["class_name"]]
        args = {}
        for key, val in source["arguments"].items():
            if isinstance(val, dict) and (set(val) == {"ModelConfig"} or ("class_name" in val and "arguments" in val)):
                val = Model.from_config(val)
            args[key] = val
        return klass(**args)

    def load_state_dict(self, state_dict: Union[str, Dict], **kwargs: Any):  # type: ignore[override]
        if isinstance(state_dict, str):
            state_dict = torch.load(state_dict, map_location="cpu")
        renamed = {k.replace("_gnn", "backbone"): v for k, v in state_dict.items()}  # model.py:72-74
        return super().load_state_dict(renamed, **kwargs)

    @property
    def device(self) -> torch.device:
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")


def _all_subclasses(cls):
    out = set()
    for sub in cls.__subclasses__():
        out.add(sub)
        out |= _all_subclasses(sub)
    return out
