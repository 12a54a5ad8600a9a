#!/bin/bash
# One GPU trip: each test file separately (own timeout, keep going), logs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for f in "$@"; do
  name=$(basename $f .py)
  timeout -k 10 300 python -m pytest $f -m gpu -q --timeout 120 -p no:cacheprovider $PYTEST_EXTRA > gpurun_out/$name.log 2>&1
  echo "exit=$?" >> gpurun_out/$name.log
  echo "=== $name"; tail -25 gpurun_out/$name.log
done
