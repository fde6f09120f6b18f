#!/bin/sh
# oracle/make_ref.sh -- stage the REAL reference (damlobster/DotsBoxesAZ, pure Python) as the CPU arm of bench.py.
#
# TEST INFRASTRUCTURE.  Copies the reference's own modules for the self-play hot path, unmodified, from where they lie
# (/root/reference, read-only) into oracle/_ref/ -- which is git-ignored (no reference source ever enters the history)
# but NOT gpurun-ignored, so it travels to the GPU box like a built .so -- and adds the two shims the reference needs
# in this image: `async_timeout` (-> asyncio.timeout) and `pylru` (-> an OrderedDict LRU), both absent
# from the offline wheelhouse.  Only bench.py's `--impl reference` / `cpu_baseline` legs and tests/ may import it.
#
#   sh oracle/make_ref.sh [/root/reference]
set -e
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/mcts.py" ]; then
    echo "make_ref: $SRC/mcts.py not found (the reference is only present in the build container)" >&2
    exit 3
fi
rm -rf "$DST"
mkdir -p "$DST/dots_boxes" "$DST/utils"
for f in __init__.py game.py mcts.py nn.py self_play.py configuration.py; do cp "$SRC/$f" "$DST/$f"; done
for f in __init__.py dots_boxes_game.py dots_boxes_nn.py; do cp "$SRC/dots_boxes/$f" "$DST/dots_boxes/$f"; done
for f in __init__.py proxies.py utils.py; do cp "$SRC/utils/$f" "$DST/utils/$f"; done
chmod -R u+w "$DST"

cat > "$DST/async_timeout.py" <<'PY'
"""Shim (oracle/make_ref.sh): the reference imports `from async_timeout import timeout` (utils/proxies.py:5) and uses
`async with timeout(seconds)`; asyncio.timeout (Python >= 3.11) has the same contract and raises TimeoutError, which
is asyncio.TimeoutError."""
import asyncio


def timeout(delay):
    return asyncio.timeout(delay)
PY

cat > "$DST/pylru.py" <<'PY'
"""Shim (oracle/make_ref.sh): the subset of pylru the reference uses (utils/proxies.py:23-26,35-43): lrucache(size)
with `in`, item get (refreshes recency) and item set (evicts the least recently used entry beyond `size`)."""
from collections import OrderedDict


class lrucache:
    def __init__(self, size):
        self.size = size
        self._d = OrderedDict()

    def __contains__(self, key):
        return key in self._d

    def __getitem__(self, key):
        self._d.move_to_end(key)
        return self._d[key]

    def __setitem__(self, key, value):
        self._d[key] = value
        self._d.move_to_end(key)
        while len(self._d) > self.size:
            self._d.popitem(last=False)

    def __len__(self):
        return len(self._d)
PY
echo "make_ref: reference staged in $DST"
