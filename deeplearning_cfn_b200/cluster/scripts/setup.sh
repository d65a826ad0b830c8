#!/bin/bash
# Per-node environment setup (counterpart of examples/distributed-tensorflow/setup.sh, which pip-pins
# numpy/opencv/Cython and builds pycocotools).  Offline B200 image: verify the pinned stack instead of
# installing it, and build the in-tree native libraries if they are stale.
set -e
DLCFN_PY=${DLCFN_PY:-python}
$DLCFN_PY - <<'PY'
import importlib, sys
need = {"torch": "2.", "numpy": ""}
for mod, prefix in need.items():
    m = importlib.import_module(mod)
    v = getattr(m, "__version__", "?")
    if prefix and not v.startswith(prefix):
        sys.exit(f"{mod} {v} does not satisfy pinned {prefix}x")
    print(f"{mod}=={v}")
PY
$DLCFN_PY -m deeplearning_cfn_b200._build
echo "setup complete"
