#!/bin/bash
# gpu.sh [gpurun options] -- 'command' : rebuild every in-tree artefact (a stale .so would travel otherwise), then gpurun
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -v "warning\|colnorm2\|^ *|\|^ *[0-9]* |\|In file\|In static\|^\s*\^" | head -5
exec /usr/local/graft/bin/gpurun "$@"
