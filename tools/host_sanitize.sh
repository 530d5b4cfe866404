#!/bin/bash
# The host layer (csrc/host + pybind) under AddressSanitizer + UndefinedBehaviorSanitizer:
# an instrumented copy of the extension and of libllo.so is built in a scratch copy of the
# tree, the CPU suite and the C++ callers run against it.  No GPU needed (the CUDA library
# is the shipped one).   usage: tools/host_sanitize.sh [scratch dir]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=${1:-/tmp/llo_host_sanitize}
rm -rf "$WORK" && mkdir -p "$WORK"
(cd "$ROOT" && git archive HEAD) | tar -x -C "$WORK"
cp -a "$ROOT/cortenn_b200/libllo_cuda.so" "$WORK/cortenn_b200/"
[ -d "$ROOT/oracle/_build" ] && cp -a "$ROOT/oracle/_build" "$WORK/oracle/"
cd "$WORK"
sed -i 's/"-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden",/"-O1", "-g", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",/' cortenn_b200/build.py
sed -i 's/\["-L", HERE, "-lllo_cuda", "-lstdc++", "-lm",/["-fsanitize=address,undefined", "-L", HERE, "-lllo_cuda", "-lstdc++", "-lm",/g' cortenn_b200/build.py
python - <<PY
import importlib.util
spec = importlib.util.spec_from_file_location("b", "$WORK/cortenn_b200/build.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
m.build_host()
PY
[ -d oracle/_build ] || python oracle/build_oracle.py
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
  python -m pytest tests/ -q -s -m "not gpu" -p no:cacheprovider \
  --deselect tests/test_host_cpp_api.py --deselect tests/test_capi_symbols.py::test_header_is_plain_c_and_links \
  > sanitize_pytest.log 2>&1 || true
echo "pytest: $(tail -1 sanitize_pytest.log)"
echo "UBSan reports: $(grep -c 'runtime error' sanitize_pytest.log)   ASan reports: $(grep -c AddressSanitizer sanitize_pytest.log)"
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Icortenn_b200/csrc/host -Iinclude tests/cpp/host_api.cpp \
  -Lcortenn_b200 -lllo -lllo_cuda -Wl,-rpath,"$WORK/cortenn_b200" -o host_api_asan
ASAN_OPTIONS=detect_leaks=1 ./host_api_asan > host_api.log 2>&1; echo "host_api (ASan+LSan+UBSan): rc=$? $(grep -c -E 'runtime error|Sanitizer' host_api.log) reports"
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -Icortenn_b200/csrc/host tests/cpp/shear_prune.cpp -o shear_asan
ASAN_OPTIONS=detect_leaks=1 ./shear_asan > shear.log 2>&1; echo "shear_prune (ASan+LSan+UBSan): rc=$? $(grep -c -E 'runtime error|Sanitizer' shear.log) reports"
