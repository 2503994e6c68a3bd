#!/bin/sh
# Builds the oracle restatement (liboracle.so) and, when /root/reference exists, the reference driver.
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -mavx2 -mfma -shared -fPIC -o liboracle.so fs_oracle.cpp
if [ -d /root/reference/src ]; then
    make -j"$(nproc)" >/dev/null
fi
