#!/bin/sh -e
# The reference's example/ proteins (as databases written by the reference's createdb: tests/golden/example) through foldseek-b200 on a B200.
#   sh examples/search_example.sh [workdir]
# 1. easy-search: d1asha_ against all 30 chains -> aln.m8 (what `foldseek easy-search example/d1asha_ example/ aln.m8 tmp` writes)
# 2. the two modules the way data/structuresearch.sh chains them
# 3. a resident server: stage the target once, serve several searches
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
BIN="$ROOT/foldseek_b200/bin/foldseek-b200"
W="${1:-$(mktemp -d /tmp/fsb200_example.XXXXXX)}"
cp -r "$ROOT/tests/golden/example/." "$W/"
cd "$W"

"$BIN" easy-search q db aln.m8 tmp
echo "easy-search: $(wc -l < aln.m8) rows; identical to the reference's output: $(cmp -s aln.m8 ref_aln_q.m8 && echo yes || echo NO)"

"$BIN" prefilter q_ss db_ss pref -s 9.5 --max-seqs 1000 --comp-bias-corr-scale 0.15
"$BIN" structurealign q db pref aln -e 10 --comp-bias-corr-scale 0.5 -a 1
"$BIN" convertalis q db aln aln_modules.m8
echo "modules: $(wc -l < aln_modules.m8) rows"

FSB200_SOCKET_DIR="$W" "$BIN" gpuserver db --lanes 2 > server.log 2>&1 &
SERVER=$!
trap 'kill -TERM $SERVER 2>/dev/null || true' EXIT
for i in 1 2 3; do
    FSB200_SOCKET_DIR="$W" "$BIN" search db db "all_$i" tmp --gpu-server 1 | tail -1
done
kill -TERM $SERVER; wait $SERVER || true
trap - EXIT
echo "results in $W"
