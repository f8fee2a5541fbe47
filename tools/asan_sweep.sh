# Memory-safety sweep of the host parser and the kernel bodies on corrupt input, without a GPU: builds the CPU
# emulation of the kernel bodies (tests/hostemu) with AddressSanitizer + UBSan and runs it over a corpus of
# corrupted streams. usage: python tools/gen_corrupt_corpus.py 1 2000 /tmp/corpus && bash tools/asan_sweep.sh /tmp/corpus
# (round 1: 14 000 corrupt + 600 valid streams; two findings, both fixed)
set -e
SRC=j40_b200/csrc
g++ -O1 -g -std=c++17 -ffp-contract=off -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer \
    -Wno-unused-function tools/asan_emu_main.cc tests/hostemu/hostemu.cc $SRC/j40b_host.cc -o /tmp/asan_emu -lm
for f in "$1"/*.jxl; do
  out=$(/tmp/asan_emu "$f" 2>&1) || true
  if echo "$out" | grep -q "runtime error\|AddressSanitizer\|Segmentation"; then echo "=== $f"; echo "$out" | grep -E "runtime error|ERROR: AddressSanitizer|#[0-3] " | head -8; fi
done
