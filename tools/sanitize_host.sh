#!/bin/bash
# Host-side C++ of the library under sanitizers (CPU only):
#   csrc/hostpack.cpp  ThreadSanitizer, 200 runs of 4 jobs over a 6-thread pool (int32 / int64,
#                      ragged tails, an id that does not fit)   [per-ISA clones off: ifunc
#                      resolvers run before the TSan runtime is up]
#   csrc/wire.cpp      AddressSanitizer + UBSan, 20,000 parses of truncated / bit-flipped windows
#                      of 200 real frames on exact-size heap copies, each followed by an emit
#                      into a buffer that is sometimes too small
set -e
cd "$(dirname "$0")/sanitize"
T=$(mktemp -d)
g++ -O1 -g -std=c++17 -fsanitize=thread -pthread '-Dtarget_clones(...)=unused' -o $T/tsan hostpack_tsan.cpp ../../deeprecsys_b200/csrc/hostpack.cpp
echo "== hostpack.cpp under ThreadSanitizer"; DRS_HOSTPACK_NT=0 $T/tsan; echo "rc=$?"
python make_frames.py $T/frames.bin
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -I ../../include -o $T/asan wire_asan.cpp ../../deeprecsys_b200/csrc/wire.cpp
echo "== wire.cpp under AddressSanitizer + UBSan"; (cd $T && ./asan); echo "rc=$?"
rm -rf $T
