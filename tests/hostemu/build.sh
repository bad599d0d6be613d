#!/bin/sh
# Build the CPU emulation of the Ops (tests only).  Output: tests/hostemu/_build/libdicp_hostemu.so (git-ignored).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_build"
# compiled next to the target and renamed: processes that build at the same time (the two ranks of the gloo tests) never load a
# half-written library
tmp="$here/_build/libdicp_hostemu.so.$$"
g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -I/usr/local/cuda/include -o "$tmp" "$here/hostemu.cpp"
mv -f "$tmp" "$here/_build/libdicp_hostemu.so"
