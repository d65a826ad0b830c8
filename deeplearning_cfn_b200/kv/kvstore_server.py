"""Server / scheduler side of a parameter-server job.

Reference: `MX/python/mxnet/kvstore_server.py` - on `import mxnet`, a process whose DMLC_ROLE is not
`worker` creates a `dist` kvstore, runs the server loop and exits (:75-85); command 0 (`kController`) carries
a pickled optimizer which the server installs as its updater (:41-62).  Here the hook fires when
`deeplearning_cfn_b200.kv` is imported (or `init_server_module()` is called): servers run the native request
loop of csrc/rt (reduce + optimizer in C++), falling back to a python `Updater` only for optimizers the C++
server does not implement.
"""
from __future__ import annotations

import ctypes
import os
import pickle
import sys

import numpy as np
import torch

from ..utils import rt
from . import optimizer as opt


class KVStoreServer:
    def __init__(self):
        self.L = rt.lib()
        self.updater = None
        self._cb_ctl = rt.CONTROLLER_FN(self._controller)
        self._cb_upd = rt.UPDATER_FN(self._update)

    def _controller(self, head, body, length, _user):
        if head == rt.CMD_CONTROLLER and length:
            raw = ctypes.string_at(body, length)
            try:
                optimizer = pickle.loads(raw)
            except Exception as e:      # a server that cannot build its updater must say so: workers would hang
                print(f"[dlcfn server] cannot unpickle the optimizer sent by the workers: {e!r}", file=sys.stderr, flush=True)
                return
            if isinstance(optimizer, opt.Optimizer):
                self.updater = opt.get_updater(optimizer)
                self.L.dlcfn_kv_server_set_updater(self._cb_upd, None)

    def _update(self, key, grad, weight, n, _user):
        g = torch.from_numpy(np.ctypeslib.as_array(grad, shape=(n,)))
        w = torch.from_numpy(np.ctypeslib.as_array(weight, shape=(n,)))
        self.updater(key, g, w)

    def run(self) -> None:
        from ..utils.cpu import share_cpu_threads
        share_cpu_threads()                 # python updaters on a server use torch ops: split the cores with local peers
        if self.L.dlcfn_ps_start() != 0:
            raise RuntimeError("parameter-server node could not start")
        role = self.L.dlcfn_ps_role()
        if role == rt.ROLE_SERVER:
            self.L.dlcfn_kv_server_set_controller(self._cb_ctl, None)
        if not self.L.dlcfn_ps_is_recovery():
            self.L.dlcfn_ps_barrier(rt.GROUP_ALL)
        if role == rt.ROLE_SERVER:
            self.L.dlcfn_kv_server_run()          # returns after kStopServer
        self.L.dlcfn_ps_finalize(1)


def init_server_module() -> None:
    """If this process is a scheduler or a server of a PS job: serve, then exit (never returns)."""
    role = os.environ.get("DMLC_ROLE", "worker")
    if role in ("server", "scheduler") and "DMLC_PS_ROOT_URI" in os.environ:
        KVStoreServer().run()
        sys.exit(0)
