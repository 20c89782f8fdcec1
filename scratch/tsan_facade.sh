#!/bin/bash
# Host threading of the facade (HDLSource receive/consumer threads, INSSource, HDLManager cache and
# buffers) under ThreadSanitizer: builds the driver + facade sources with -fsanitize=thread and runs
# the CPU facade tests with it.  Reports land in /tmp/tsan/log.* (none = clean).
cd "$(dirname "$0")/.."
mkdir -p /tmp/tsan && rm -f /tmp/tsan/log*
g++ -O1 -g -std=c++17 -ffp-contract=off -fsanitize=thread -fPIE -pie -o /tmp/tsan/facade_driver \
  tests/cpp/facade_driver.cpp veloslam_b200/cpp/*.cpp -I veloslam_b200/cpp \
  -L veloslam_b200 -lveloslam_b200 -lpthread -Wl,-rpath,$PWD/veloslam_b200 || exit 1
VS_TEST_DRIVER=/tmp/tsan/facade_driver TSAN_OPTIONS="log_path=/tmp/tsan/log exitcode=0" \
  python -m pytest tests/test_facade_host.py -q
ls /tmp/tsan/log* 2>/dev/null && grep -c "WARNING: ThreadSanitizer" /tmp/tsan/log* || echo "ThreadSanitizer: no reports"
