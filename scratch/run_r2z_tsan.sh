#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out /tmp/tsan && rm -f /tmp/tsan/log*
g++ -O1 -g -std=c++17 -ffp-contract=off -fsanitize=thread -fPIE -pie -o /tmp/tsan/facade_driver \
  tests/cpp/facade_driver.cpp veloslam_b200/cpp/*.cpp -I veloslam_b200/cpp \
  -L veloslam_b200 -lveloslam_b200 -lpthread -Wl,-rpath,$PWD/veloslam_b200 || exit 1
export VS_TEST_DRIVER=/tmp/tsan/facade_driver TSAN_OPTIONS="log_path=/tmp/tsan/log exitcode=0 report_signal_unsafe=0"
timeout 600 python -m pytest tests/test_gpu_facade.py -q -m gpu -k "online_udp or spread or pipelined or streaming" 2>&1 | tail -3
ls /tmp/tsan/ | head
cat /tmp/tsan/log* 2>/dev/null | grep -E "WARNING: ThreadSanitizer" | sort | uniq -c | head
cat /tmp/tsan/log* 2>/dev/null | head -150 > gpurun_out/r2z_tsan_gpu.txt
wc -l gpurun_out/r2z_tsan_gpu.txt
