"""rtfs_net_b200 -- B200-native (sm_100a) implementation of the RTFS-Net model-forward hot path.

Mirrors the reference's `src/models` entry points (src/models/__init__.py:9-42): `AVNet`,
`get(name)` (case-insensitive) and `register_model`.
"""
from .nn import AVNet

__all__ = ["AVNet", "get", "register_model"]

_REGISTRY = {"AVNet": AVNet}


def register_model(custom_model):
    """Register a custom model, gettable with `get` (src/models/__init__.py:15-24)."""
    name = custom_model.__name__
    if name in _REGISTRY or name.lower() in {k.lower() for k in _REGISTRY}:
        raise ValueError(f"Model {name} already exists. Choose another name.")
    _REGISTRY[name] = custom_model


def get(identifier):
    """Model class from a string, case-insensitive (src/models/__init__.py:27-42)."""
    if isinstance(identifier, str):
        cls = {k.lower(): v for k, v in _REGISTRY.items()}.get(identifier.lower())
        if cls is not None:
            return cls
    raise ValueError(f"Could not interpret model name : {str(identifier)}")
