#!/bin/sh
# Pins the oracle against the Go reference on a machine that has Go (this image has none):
#   oracle/goref/run.sh /path/to/a/checkout/of/tanghaibao/allhic
# Copies the dump test next to the reference's sources, runs it with the reference's own
# modules, leaves tests/golden/go_reference_dump.json in THIS repo, removes the copy.
set -e
REF=${1:?usage: run.sh <checkout of github.com/tanghaibao/allhic>}
HERE=$(cd "$(dirname "$0")" && pwd)
cp "$HERE/ahx_dump_test.go" "$REF/ahx_dump_test.go"
trap 'rm -f "$REF/ahx_dump_test.go"' EXIT
(cd "$REF" && AHX_GOLDEN="$HERE/../../tests/golden" go test -run TestAhxDump -count=1 -v .)
echo "now run: python -m pytest tests/test_oracle.py -k go_reference -q"
