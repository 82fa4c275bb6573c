#!/bin/bash
# Retry wrapper for gpurun: a busy pod answers with exit code 3 (nothing charged); wait and try again.
# Usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt: busy, sleeping 90 s" >&2
  sleep 90
done
exit 3
