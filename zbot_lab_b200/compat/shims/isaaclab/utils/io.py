"""``isaaclab.utils.io.dump_yaml / dump_pickle`` (train.py:91, 199-202)."""
import os
import pickle

import yaml


def _to_plain(data):
    if hasattr(data, "to_dict"):
        data = data.to_dict()
    if isinstance(data, dict):
        return {str(k): _to_plain(v) for k, v in data.items()}
    if isinstance(data, (list, tuple)):
        return [_to_plain(v) for v in data]
    if isinstance(data, (int, float, str, bool)) or data is None:
        return data
    return str(data)


def dump_yaml(filename: str, data, sort_keys: bool = False):
    if not filename.endswith("yaml"):
        filename += ".yaml"
    os.makedirs(os.path.dirname(filename), exist_ok=True)
    with open(filename, "w") as f:
        yaml.safe_dump(_to_plain(data), f, default_flow_style=False, sort_keys=sort_keys)


def dump_pickle(filename: str, data):
    if not filename.endswith("pkl"):
        filename += ".pkl"
    os.makedirs(os.path.dirname(filename), exist_ok=True)
    with open(filename, "wb") as f:
        pickle.dump(_to_plain(data), f)
