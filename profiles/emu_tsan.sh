#!/bin/bash
# ThreadSanitizer run of the emulated kernels: the emulator maps every CUDA thread to a host thread, so a
# missing __syncthreads() / __syncwarp() shows up as a data race.  Prints the number of reports (expected: 0).
#   bash profiles/emu_tsan.sh
TSAN=$(gcc -print-file-name=libtsan.so)
LD_PRELOAD=$TSAN TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 history_size=4" python profiles/emu_tsan_case.py > /tmp/fgs_tsan.log 2>&1
grep -E "done" /tmp/fgs_tsan.log
echo "ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' /tmp/fgs_tsan.log)"
