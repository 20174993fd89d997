mkdir -p gpurun_out
for d in 0 8 16 1 9 17 32; do echo "== VKOCR_DEBUG_SKIP_TMA=$d"; VKOCR_DEBUG_SKIP_TMA=$d python tools/kbench.py mlp 2>&1 | grep -E "M819200 K96 N384|M51200 K384 N1536|M819200 K384 N96" | grep -E "mlp1 (fwd|eval)|mlp2 (fwd|dgrad)"; done > gpurun_out/x9_gemm_debug.log 2>&1
cat gpurun_out/x9_gemm_debug.log
