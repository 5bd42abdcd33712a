#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/epi_ncu2.py > $O/r02zb_epi_ncu2_plain.log 2>&1; echo "plain rc=$?"; tail -3 $O/r02zb_epi_ncu2_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -f -o $O/r02zb_ncu_epi2 python scripts/epi_ncu2.py > $O/r02zb_ncu_epi2.log 2>&1; echo "ncu rc=$?"; tail -3 $O/r02zb_ncu_epi2.log
ls -la $O/r02zb_ncu_epi2.ncu-rep
