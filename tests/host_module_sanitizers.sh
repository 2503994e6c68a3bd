#!/bin/sh -e
# Builds the host module (cli_main.cpp + cli_*.h) with AddressSanitizer + UndefinedBehaviorSanitizer on the oracle-backed stand-in for the C ABI
# and drives the main paths on the example databases: modules with lanes, alternative alignments, target split, search, easy-search, and a
# gpuserver session (leak check at exit).  Test infrastructure; prints sanitizer reports, if any.
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W="$(mktemp -d /tmp/fsb_asan.XXXXXX)"
g++ -O1 -g -std=c++17 -fopenmp -mavx2 -mfma -fsanitize=address,undefined -fno-omit-frame-pointer -o "$W/mock_asan" \
    "$ROOT/foldseek_b200/csrc/cli_main.cpp" "$ROOT/oracle/mock_fsb200.cpp" "$ROOT/oracle/fs_oracle.cpp" -I"$ROOT/include" -lpthread -ldl
cp -r "$ROOT/tests/golden/example/." "$W/"
cd "$W"
export FSB200_DATA="$ROOT/foldseek_b200/data" FSB200_SOCKET_DIR="$W" UBSAN_OPTIONS=print_stacktrace=1
M="$W/mock_asan"
PF="$(grep '^prefilter' params.txt | cut -d' ' -f2-)"
AL="$(grep '^structurealign' params.txt | cut -d' ' -f2-)"
{
    ASAN_OPTIONS=detect_leaks=0 $M prefilter db_ss db_ss p1 $PF --lanes 2
    ASAN_OPTIONS=detect_leaks=0 $M structurealign db db ref_pref_all a1 $AL --alt-ali 2 -a 1 --lanes 2
    ASAN_OPTIONS=detect_leaks=0 $M search q db s1 tmp --split 3 --split-mode 0
    ASAN_OPTIONS=detect_leaks=0 $M easy-search q db e.m8 tmpx
    $M gpuserver db --lanes 2 &
    SP=$!
    $M prefilter q_ss db_ss sp $PF --gpu-server 1
    $M search q db ss tmp --gpu-server 1
    $M search q db ss2 tmp --max-seqs 5 --gpu-server 1
    $M structurealign q db nonexist x --gpu-server 1 || true
    $M search q db ss3 tmp --alt-ali 1 --gpu-server 1
    kill -TERM $SP; wait $SP
} > run.log 2>&1
if grep -n "ERROR: AddressSanitizer\|runtime error\|LeakSanitizer" run.log; then echo "sanitizer reports above ($W/run.log)"; exit 1; fi
cmp e.m8 ref_aln_q.m8
echo "host module: no AddressSanitizer / UBSan / leak reports"
# ThreadSanitizer on the multi-context paths (one host thread per GPU and lane; built without OpenMP: libgomp is not instrumented)
g++ -O1 -g -std=c++17 -mavx2 -mfma -fsanitize=thread -o "$W/mock_tsan" "$ROOT/foldseek_b200/csrc/cli_main.cpp" "$ROOT/oracle/mock_fsb200.cpp" \
    "$ROOT/oracle/fs_oracle.cpp" -I"$ROOT/include" -lpthread -ldl 2>/dev/null
T="$W/mock_tsan"
{
    $T prefilter db_ss db_ss tp $PF --gpu-list 0,0 --lanes 2
    $T structurealign db db ref_pref_all ta $AL --alt-ali 1 --gpu-list 0,0 --lanes 2
    $T search db db ts tmp --lanes 3
    $T prefilter db_ss db_ss tsp $PF --split 3 --split-mode 0 --max-seqs 10 --gpu-list 0,0 --lanes 2     # target ranges sharded over the GPUs
    $T gpuserver db --lanes 2 &
    SP=$!
    $T search db db t1 tmp --gpu-server 1
    $T prefilter db_ss db_ss t2 $PF --gpu-server 1
    kill -TERM $SP; wait $SP
} > tsan.log 2>&1
if grep -n "WARNING: ThreadSanitizer" tsan.log; then echo "ThreadSanitizer reports above ($W/tsan.log)"; exit 1; fi
echo "host module: no ThreadSanitizer reports"
rm -rf "$W"
