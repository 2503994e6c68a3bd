#!/bin/sh
# Runs the CLI-driven -m gpu tests (tests/test_cli_gpu.py, test_example_gpu.py, test_workflow_gpu.py, test_zz_host_features_gpu.py) against the
# oracle-backed stand-in for the C ABI (oracle/_mock/foldseek-b200-mock) instead of the CUDA build: a dress rehearsal of their host-module
# expectations (flags, messages, file handling, goldens) on a machine without a GPU.  Test infrastructure; the real run is `pytest -m gpu` on a B200.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
D="$(mktemp -d /tmp/fsb_rehearsal.XXXXXX)"
mkdir -p "$D/tests"
cp "$ROOT/tests/conftest.py" "$D/tests/"
for f in test_cli_gpu test_example_gpu test_workflow_gpu test_zz_host_features_gpu; do
    sed -e "s#^CLI = .*#CLI = \"$ROOT/oracle/_mock/foldseek-b200-mock\"#" -e "s#^pytestmark = pytest.mark.gpu##" \
        -e "s#^ROOT = .*#ROOT = \"$ROOT\"#" "$ROOT/tests/$f.py" > "$D/tests/$f.py"
done
cd "$D"
FSB200_DATA="$ROOT/foldseek_b200/data" PYTHONPATH="$ROOT" python -m pytest tests -q -x "$@"
rm -rf "$D"
