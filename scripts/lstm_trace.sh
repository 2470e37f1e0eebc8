#!/bin/bash
for M in ${MS:-512 1024}; do echo "== M=$M"; MARLC_LIB=$PWD/marlclassification_b200/libmarlc_trace.so python scripts/lstm_once.py $M 2>&1 | grep -E "tc trace|cell:" | tail -4; done
