#!/bin/bash
# build everything in-tree from any cwd; exit non-zero on failure
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
