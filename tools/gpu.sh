#!/bin/bash
# Build the library, check that the C-ABI loads, THEN snapshot to the GPU box: a stale .so next to new Python bindings wastes a call.
# usage: tools/gpu.sh [--gpus N] [--timeout S] -- '<command>'
set -e
cd "$(dirname "$0")/.."
python -m mupopp_b200.build > /dev/null
python -c "from mupopp_b200 import _lib; _lib.load()"
exec /usr/local/graft/bin/gpurun "$@"
