#!/bin/bash
# Rebuild the library (a stale .so must never travel), then hand the command to gpurun.
set -e
cd "$(dirname "$0")/.."
./inverseproblemwithdiffusionmodel_b200/csrc/build.sh > /tmp/ipdm_build.log 2>&1 || { cat /tmp/ipdm_build.log; exit 1; }
exec /usr/local/graft/bin/gpurun "$@"
