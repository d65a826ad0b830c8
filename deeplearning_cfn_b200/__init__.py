"""deeplearning_cfn_b200 - a B200-native distributed-training launcher + runtime with the capabilities of
awslabs/deeplearning-cfn (cluster provisioning / rendezvous / hostfile + env contract, launchers, KVStore
parameter server, Horovod-style all-reduce, example workloads), rebuilt for one NVSwitch-connected box of
Blackwell GPUs: PyTorch for the substrate, hand-written sm_100a kernels for the communication-bound hot
paths, NCCL kept as the baseline."""
__version__ = "0.1.0"

# Same contract as `import mxnet` in the reference (MX/python/mxnet/kvstore_server.py:75-85): a process started by the
# launcher with DMLC_ROLE=server|scheduler turns into that node when the PACKAGE is imported - whatever the user script
# goes on to import - and exits when the job ends.  Importing `.kv` runs `init_server_module()`.
import os as _os

if _os.environ.get("DMLC_ROLE") in ("server", "scheduler") and "DMLC_PS_ROOT_URI" in _os.environ \
        and _os.environ.get("DLCFN_NO_SERVER_HOOK") != "1":
    from . import kv as _kv  # noqa: F401  (serves, then sys.exit(0))
