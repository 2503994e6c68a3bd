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
fi
