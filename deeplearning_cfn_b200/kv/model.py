"""Training-loop helpers with the names and behaviour of the reference's `mxnet/model.py` (Y3):

  _create_kvstore(kvstore, num_device, arg_params)      (model.py:77-114)  - which store, and update_on_kvstore
  _initialize_kvstore(kvstore, param_arrays, arg_params, param_names, update_on_kvstore)   (:116-123)
  _update_params_on_kvstore(param_arrays, grad_arrays, kvstore, param_names)               (:145-155)
  _update_params_on_kvstore_nccl(...)                                                      (:125-143)
  _update_params(param_arrays, grad_arrays, updater, num_device, kvstore, param_names)     (:157-176)
  save_checkpoint(prefix, epoch, symbol, arg_params, aux_params) / load_checkpoint(prefix, epoch)   (:383-440)

`param_arrays[i]` / `grad_arrays[i]` are lists with one tensor per device, as in the reference's executor group.
"""
from __future__ import annotations

import json
import logging
import os

import torch

from . import kvstore as kvs

BIG_ARRAY_FOR_LOCAL_UPDATE = 1024 * 1024 * 16     # model.py:100-108


def _create_kvstore(kvstore, num_device: int, arg_params: dict):
    """Returns (kv, update_on_kvstore).  A string names a store type; with a single device and a non-dist type no
    store is needed; `local` stores update on the kvstore unless some parameter is larger than 16 Mi elements."""
    update_on_kvstore = True
    if kvstore is None:
        kv = None
    elif isinstance(kvstore, kvs.KVStore):
        kv = kvstore
    elif isinstance(kvstore, str):
        if num_device == 1 and "dist" not in kvstore:
            kv = None
        else:
            kv = kvs.create(kvstore)
            if kvstore == "local":
                max_size = max((int(p.numel()) for p in arg_params.values()), default=0)
                if max_size > BIG_ARRAY_FOR_LOCAL_UPDATE:
                    update_on_kvstore = False
    else:
        raise TypeError("kvstore must be KVStore, str or None")
    if kv is None:
        update_on_kvstore = False
    return kv, update_on_kvstore


def _initialize_kvstore(kvstore, param_arrays, arg_params: dict, param_names, update_on_kvstore: bool) -> None:
    for idx, param_on_devs in enumerate(param_arrays):
        name = param_names[idx]
        kvstore.init(name, arg_params[name])
        if update_on_kvstore:
            kvstore.pull(name, out=param_on_devs, priority=-idx)


def _update_params_on_kvstore(param_arrays, grad_arrays, kvstore, param_names) -> None:
    """push the gradients, pull the updated weights - per parameter, front layers first (priority = -index)."""
    for index, (arg_list, grad_list) in enumerate(zip(param_arrays, grad_arrays)):
        if grad_list[0] is None:
            continue
        name = param_names[index]
        kvstore.push(name, grad_list, priority=-index)
        kvstore.pull(name, out=arg_list, priority=-index)


def _update_params_on_kvstore_nccl(param_arrays, grad_arrays, kvstore, param_names) -> None:
    valid = [i for i, g in enumerate(grad_arrays) if g[0] is not None]
    size = int(os.environ.get("MXNET_UPDATE_AGGREGATION_SIZE", "16"))
    start = 0
    while start < len(valid):
        idx = valid[start:start + size]
        names = [param_names[i] for i in idx]
        kvstore.push(names, [grad_arrays[i] for i in idx], priority=-start)
        kvstore.pull(names, out=[param_arrays[i] for i in idx], priority=-start)
        start += size


def _update_params(param_arrays, grad_arrays, updater, num_device: int, kvstore=None, param_names=None) -> None:
    """Aggregate the gradients through the store (if any), then run the local updater once per device copy; the
    updater key is index * num_device + k so that every device copy owns its optimizer state (model.py:169-176)."""
    for index, (arg_list, grad_list) in enumerate(zip(param_arrays, grad_arrays)):
        if grad_list[0] is None:
            continue
        if kvstore is not None:
            name = param_names[index]
            kvstore.push(name, grad_list, priority=-index)
            kvstore.pull(name, out=grad_list, priority=-index)
        for k, (w, g) in enumerate(zip(arg_list, grad_list)):
            updater(index * num_device + k, g, w)


def save_checkpoint(prefix: str, epoch: int, symbol, arg_params: dict, aux_params: dict) -> None:
    """`prefix-symbol.json` + `prefix-%04d.params` with `arg:` / `aux:` key prefixes."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    if symbol is not None:
        desc = symbol if isinstance(symbol, dict) else {"class": type(symbol).__name__, "repr": repr(symbol)[:20000]}
        with open(f"{prefix}-symbol.json", "w") as f:
            json.dump(desc, f)
    sd = {f"arg:{k}": v.detach().cpu() for k, v in arg_params.items()}
    sd.update({f"aux:{k}": v.detach().cpu() for k, v in aux_params.items()})
    name = f"{prefix}-{epoch:04d}.params"
    torch.save(sd, name)
    logging.info('Saved checkpoint to "%s"', name)


def load_checkpoint(prefix: str, epoch: int):
    """Returns (symbol description, arg_params, aux_params)."""
    symbol = None
    if os.path.exists(f"{prefix}-symbol.json"):
        with open(f"{prefix}-symbol.json") as f:
            symbol = json.load(f)
    raw = torch.load(f"{prefix}-{epoch:04d}.params", map_location="cpu")
    arg = {k[4:]: v for k, v in raw.items() if k.startswith("arg:")}
    aux = {k[4:]: v for k, v in raw.items() if k.startswith("aux:")}
    return symbol, arg, aux
