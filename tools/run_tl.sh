#!/bin/bash
# usage: tools/run_tl.sh "ENV=.." ...  -> timeline per setting
for cfg in "$@"; do
  echo "=== [$cfg]"
  env $cfg python tools/timeline.py 2>&1 | grep -v "^# kernel"
done
