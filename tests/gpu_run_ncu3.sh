#!/bin/bash
# developer tool (GPU box): ncu --set full of the translated streaming kernel with TRAM on cfg3 (16 384 instances, ring of 8 192, one 1 024-period launch)
T=${1:-r02ncu3}; O=gpurun_out; mkdir -p $O
export FX8010_TRANSLATE=2
python tests/probe_cfg.py cfg3 16384 1024 4 8192 > $O/${T}_probe3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fx_translated_sl -s 2 -c 1 -o $O/${T}_ncu_cfg3 python tests/probe_cfg.py cfg3 16384 1024 4 8192 > $O/${T}_ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"; cat $O/${T}_probe3.log; ls -la $O/${T}_ncu_cfg3.ncu-rep
