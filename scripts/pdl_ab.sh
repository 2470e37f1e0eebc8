#!/bin/bash
# A/B of the programmatic-dependent-launch edges (MARLC_PDL_MASK / MARLC_PDL_TRIG) on the phase tables.
OUT=gpurun_out/r3
mkdir -p $OUT
run() {  # tag, env...
  tag=$1; shift
  for wl in "c4 32" "c2 8"; do
    set -- "$@"
    f=$OUT/ab_${tag}_$(echo $wl | tr ' ' '_').txt
    env "$@" timeout 120 python scripts/phase_times.py $wl > $f 2>&1
    echo "$tag $wl: $(grep -E '^forward|^backward|^total' $f | awk '{printf "%s %s  ", $1, $2}')"
  done
}
MARLC_LIB=$PWD/marlclassification_b200/libmarlc_trace.so MARLC_PDL=0 python scripts/cnn_wide_trace.py 32 2>&1 | tail -3
run off MARLC_PDL=0
run all_top MARLC_PDL_MASK=127 MARLC_PDL_TRIG=1
run all_exit MARLC_PDL_MASK=127 MARLC_PDL_TRIG=0
run fwd_top MARLC_PDL_MASK=15 MARLC_PDL_TRIG=1
run fwd_nolstm MARLC_PDL_MASK=13 MARLC_PDL_TRIG=1
run pre_only MARLC_PDL_MASK=1 MARLC_PDL_TRIG=1
run fwd_bpre MARLC_PDL_MASK=29 MARLC_PDL_TRIG=1
