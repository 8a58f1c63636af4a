#!/bin/bash
# CPU sanitizer pass: the fp64 oracle and the host emulation of the kernel bodies built with
# -fsanitize=address,undefined, the emulation / slab / env tests run against them.  Restores the normal builds afterwards.
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
cp tests/emu/libfishgym_emu.so $T/emu.bak; cp oracle/libfishgym_oracle.so $T/oracle.bak
trap 'cp $T/emu.bak tests/emu/libfishgym_emu.so; cp $T/oracle.bak oracle/libfishgym_oracle.so' EXIT
SAN="-O1 -g -march=x86-64-v3 -fPIC -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -shared"
/usr/bin/g++ $SAN -ffp-contract=off -x c++ -o tests/emu/libfishgym_emu.so tests/emu/fishgym_emu.cpp
/usr/bin/g++ $SAN -fopenmp -o oracle/libfishgym_oracle.so oracle/fg_oracle.cpp
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
  OMP_NUM_THREADS=4 python -m pytest tests/test_emu_parity.py tests/test_slabs.py tests/test_env.py tests/test_random_cases.py tests/test_multi_direct_forcing.py tests/test_solid_force.py -q -m "not gpu" -k "not gloo" 2>&1 | tee $T/log | tail -3
# ... and once more with queued streams and emulated graphs (the scheduler and the graph recorder are test code, too)
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
  OMP_NUM_THREADS=4 FG_EMU_SCHED=rand:3 FG_EMU_GRAPHS=1 python -m pytest tests/test_emu_parity.py tests/test_random_cases.py tests/test_multi_direct_forcing.py tests/test_solid_force.py -q -m "not gpu" -k "not 16_bit" 2>&1 | tee -a $T/log | tail -3
if grep -qE "runtime error|AddressSanitizer" $T/log; then echo "SANITIZER FINDINGS"; grep -E "runtime error|AddressSanitizer" $T/log | head; exit 1; fi
echo "sanitizers: clean"
