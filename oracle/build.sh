#!/bin/sh
# Builds the oracle restatement (liboracle.so) and, when /root/reference exists, the reference driver.
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -mavx2 -mfma -shared -fPIC -o liboracle.so fs_oracle.cpp
# the host module on top of the oracle-backed stand-in for the C ABI (test infrastructure, see the header of mock_fsb200.cpp):
# lets tests/test_host_module_cpu.py compare foldseek-b200's host-side logic with the reference's modules without a GPU
mkdir -p _mock
g++ -O2 -std=c++17 -fopenmp -o _mock/foldseek-b200-mock ../foldseek_b200/csrc/cli_main.cpp mock_fsb200.cpp -I../include -L. -loracle -lpthread -ldl '-Wl,-rpath,$ORIGIN/..'
if [ -d /root/reference/src ]; then
    make -j"$(nproc)" >/dev/null
    # the reference's search workflow script for the workflow-level tests: a build artefact next to refseek (oracle/_ref/ is git-ignored and
    # travels to the GPU box; the 56 MB of generated resource headers in _ref/gen do not)
    cp -f /root/reference/data/structuresearch.sh _ref/structuresearch.sh && chmod u+w _ref/structuresearch.sh
fi
