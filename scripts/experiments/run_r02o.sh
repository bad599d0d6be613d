for K in 64; do
for sh in "" "128,8" "128,9" "128,10" "128,12" "128,7" "128,6"; do
  echo "K=$K shape=$sh" >> gpurun_out/r02o_closure.txt
  DICP_CC_SHAPE=$sh python scripts/bench_batched_closure.py --K $K >> gpurun_out/r02o_closure.txt 2>&1
done; done
for K in 8 16 24 32; do for sh in "128,8" "128,16" "128,12"; do
  echo "K=$K shape=$sh" >> gpurun_out/r02o_closure.txt
  DICP_CC_SHAPE=$sh python scripts/bench_batched_closure.py --K $K >> gpurun_out/r02o_closure.txt 2>&1
done; done
grep -o "shape=.*\|\"K\": [0-9]*\|closure_ms_median\": [0-9.]*\|Error.*" gpurun_out/r02o_closure.txt | paste - - - 
