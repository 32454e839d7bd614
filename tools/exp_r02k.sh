mkdir -p gpurun_out
ONLY=k_pll_core,k_pll_fix_par,k_front,k_agc_core,k_gardner,k_bits
PDT_DEBUG_SKIP=$ONLY python tools/timeline_inflight.py --inflight 8 > gpurun_out/r02k_tl_onlyacq_if8.txt 2>&1
PDT_DEBUG_SKIP=$ONLY python tools/timeline_inflight.py --inflight 2 > gpurun_out/r02k_tl_onlyacq_if2.txt 2>&1
python tools/timeline_inflight.py --inflight 8 > gpurun_out/r02k_tl_full_if8.txt 2>&1
grep -A3 "stream 99" gpurun_out/r02k_tl_onlyacq_if8.txt; grep -A3 "stream 99" gpurun_out/r02k_tl_onlyacq_if2.txt; grep -A14 "stream 99" gpurun_out/r02k_tl_full_if8.txt
