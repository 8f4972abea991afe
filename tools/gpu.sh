#!/bin/bash
# Rebuild the in-tree native artefacts, then run a command on the B200 box:  tools/gpu.sh [--gpus N] [--timeout S] -- '<cmd>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" | tail -1
exec /usr/local/graft/bin/gpurun "$@"
