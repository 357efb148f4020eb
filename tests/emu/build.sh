#!/bin/sh
# Build the kernel-emulation library used by the CPU-only logic tests.
set -e
here=$(cd "$(dirname "$0")" && pwd)
src="$here/../../baseband-tasks_b200/csrc"
g++ -O2 -std=c++20 -fPIC -shared -pthread -DBBT_EMULATE=1 -Wall -Wno-unknown-pragmas \
    -x c++ "$src/bbt_b200.cu" "$here/bbt_emu.cpp" -o "$here/libbbt_emu.so"
