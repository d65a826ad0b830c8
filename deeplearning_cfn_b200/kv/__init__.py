"""KVStore parameter-server API (reference: mx.kv / mxnet.gluon.Trainer / mxnet.optimizer)."""
from . import optimizer  # noqa: F401
from .kvstore import (KVStore, KVStoreDist, KVStoreDistDevice, KVStoreError, KVStoreLocal, KVStoreNCCL,  # noqa: F401
                      create)
from .kvstore_server import KVStoreServer, init_server_module  # noqa: F401
from . import trainer  # noqa: F401,E402
from .trainer import Trainer  # noqa: F401,E402
from . import initializer, metric, model  # noqa: F401,E402

# Same contract as `import mxnet` in the reference (kvstore_server.py:75-85): a process launched with
# DMLC_ROLE=server|scheduler becomes that node as soon as the kv package is imported, and exits afterwards.
init_server_module()
