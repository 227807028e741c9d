#!/bin/bash
for args in "4 4096 11008 8 0 1" "4 4096 11008 32 0 1" "4 4096 11008 8 1 1" "3 4096 11008 8 0 1" "3 4096 11008 8 1 1"; do
  echo "== seq_trace $args"; timeout 25 ./tests/perf/seq_trace $args
done
timeout 25 ./tests/perf/seq_trace_tr 4 4096 11008 8 0 1 | cut -c1-400
