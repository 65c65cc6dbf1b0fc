#!/bin/bash
cd "$(dirname "$0")"
run() { echo "== $*"; env "$@" timeout 25 ./lane_test $ARGS 2>&1 | grep -E "^ADJ  [0-9]|^ADJ  ncw"; r=${PIPESTATUS[0]}; echo "rc=$r"; return $r; }
ARGS="1024 2708 1433 10"
run NCW_ADJ=8 DBG=4
run NCW_ADJ=8 DBG=12
run NCW_ADJ=8 DBG=13
run NCW_ADJ=8 DBG=15
run NCW_ADJ=8 DBG=7
run NCW_ADJ=8 DBG=6
run NCW_ADJ=8 DBG=14
