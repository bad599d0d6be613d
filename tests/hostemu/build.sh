#!/bin/sh
# Build the CPU emulation of the Ops (tests only).  Output: tests/hostemu/_build/libdicp_hostemu.so (git-ignored).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_build"
g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -I/usr/local/cuda/include \
    -o "$here/_build/libdicp_hostemu.so" "$here/hostemu.cpp"
