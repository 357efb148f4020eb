#!/bin/sh
# Build the kernel-emulation library used by the CPU-only logic tests.
set -e
here=$(cd "$(dirname "$0")" && pwd)
src="$here/../../baseband-tasks_b200/csrc"
for unit in bbt_core bbt_fft bbt_dedisperse bbt_detect; do
  g++ -O2 -std=c++20 -fPIC -pthread -DBBT_EMULATE=1 $BBT_EMU_FLAGS -Wall -Wno-unknown-pragmas -Wno-unused-function \
      -x c++ -c "$src/$unit.cu" -o "$here/$unit.emu.o" &
done
g++ -O2 -std=c++20 -fPIC -pthread -DBBT_EMULATE=1 -Wall -c "$here/bbt_emu.cpp" \
    -o "$here/bbt_emu.emu.o" &
wait
g++ -shared -pthread "$here"/*.emu.o -o "$here/libbbt_emu.so"
rm -f "$here"/*.emu.o
